"""Multi-GPU host logic: the batch of PAIRS is sharded over one process per GPU.

Pairs are independent; the only coupling in the reference is the global mean of the loss
(common/loss_functions.py:141-144 divides by all N = B*HW elements).  Each rank therefore
runs the kernels on its shard with the GLOBAL pair count baked into the denominators
(`rgbd_loss_opts.n_pairs_global`), gradients need no communication at all, and the loss is
finished by one all-reduce(sum) of four floats.  The reference's analogue is ChainerMN's
`pure_nccl` communicator (train_rgbd.py:103-113), which it uses for weight gradients only.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .loss_functions import combine_loss_parts


class PeerComm:
    """Handle for the fused loss all-reduce over NVLink peer memory (include/rgbdgan_b200.h,
    rgbd_peer_comm_*): the finalize kernel of the loss exchanges the 4 means itself, so a sharded
    step issues no NCCL call at all.  One box, one process per GPU, world <= 16.

        comm = PeerComm(group)                      # collective: every rank of `group`
        f = LossFuncRotate(xp, peer_comm=comm)      # pairs sharded over the group
    """

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        lib = _lib.load()
        self._h = ctypes.c_void_p()
        mine = ctypes.create_string_buffer(64)
        _lib.check(lib.rgbd_peer_comm_create(self.rank, self.world, ctypes.byref(self._h), mine), "rgbd_peer_comm_create")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine.raw), group=group)
        blob = b"".join(handles)
        _lib.check(lib.rgbd_peer_comm_connect(self._h, blob), "rgbd_peer_comm_connect")
        dist.barrier(group=group)

    @property
    def handle(self):
        return self._h

    def wait(self, stream=None):
        """make `stream` (default: current) wait for the latest deferred loss exchange (defer_loss=True)"""
        st = torch.cuda.current_stream() if stream is None else stream
        _lib.check(_lib.load().rgbd_peer_comm_wait(self._h, ctypes.c_void_p(st.cuda_stream)), "rgbd_peer_comm_wait")

    def status(self, stream=None):
        """0 = every exchange so far completed; 1 = a wait for a peer ran into the time limit (a rank died or made a
        different sequence of loss calls): the loss parts of that call are undefined.  Synchronises."""
        st = torch.cuda.current_stream() if stream is None else stream
        out = ctypes.c_int(-1)
        _lib.check(_lib.load().rgbd_peer_comm_status(self._h, ctypes.c_void_p(st.cuda_stream), ctypes.byref(out)),
                   "rgbd_peer_comm_status")
        return out.value

    def close(self):
        if self._h:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)          # nobody may still be writing into a mailbox
            _lib.load().rgbd_peer_comm_destroy(self._h)
            self._h = ctypes.c_void_p()


def shard_range(n_pairs, rank, world_size):
    """contiguous, balanced split of `n_pairs` pairs: returns (start, stop) for `rank`.
    Twins stay together because pair b is (img[b], img_rot[b])."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(n_pairs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_loss(parts, lambda_geometric, group=None):
    """parts: tensor whose first four entries are this shard's means {rgb, rgb_rot, depth,
    depth_rot} computed with global denominators.  Returns (global loss, global parts)."""
    p = parts[:4].clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(p, op=dist.ReduceOp.SUM, group=group)
    return combine_loss_parts(p, lambda_geometric), p


def max_over_ranks(value, device, group=None):
    """max of a python float over the group (bench timing rule: report the slowest rank)"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
