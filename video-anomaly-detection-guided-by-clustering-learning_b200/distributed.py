"""NCCL replacement for utils/distritributed_model.py (gloo + DDP in the
reference): same function names / side effects
(``init_distributed_mode`` sets args.rank / world_size / gpu, mutes print on
non-master ranks), plus the packed all-reduce helpers the hot path needs.
One process per GPU; NCCL over NVLink 5 / NVSwitch when CUDA is present, gloo
for the CPU tests of the host logic."""
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(device_index):
    """CPUs on the NUMA node the GPU's PCIe root hangs off (sysfs local_cpulist), restricted to
    the CPUs this process may run on; empty set when the topology cannot be read."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as fh:
            return _parse_cpulist(fh.read()) & set(os.sched_getaffinity(0))
    except Exception:
        return set()


class numa_local:
    """Context manager: run the enclosed host allocations on the GPU's own NUMA node.

    Pinned host buffers are placed on the node of the allocating thread; a buffer on the far socket
    halves host->device copy bandwidth.  The reference keeps whole test videos and every training
    batch in host memory (dataset/utils_dataset.py:116-135, main_predict.py:241), so the host side
    of this path allocates its staging buffers inside ``with numa_local(gpu):``.  The previous CPU
    affinity is restored on exit (the CPU baseline keeps all cores)."""

    def __init__(self, device_index):
        self.cpus = gpu_local_cpus(device_index)
        self.prev = None

    def __enter__(self):
        if self.cpus:
            self.prev = os.sched_getaffinity(0)
            os.sched_setaffinity(0, self.cpus)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            os.sched_setaffinity(0, self.prev)
        return False


def pinned_like_local(t, device_index):
    """pinned host copy of ``t`` allocated on the NUMA node of GPU ``device_index``"""
    with numa_local(device_index):
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        out.copy_(t)
    return out


def setup_for_distributed(is_master):
    """disable printing when not in master process (utils/distritributed_model.py:23-35)"""
    import builtins as __builtin__
    builtin_print = __builtin__.print
    if getattr(builtin_print, "_vadc_wrapped", False):
        builtin_print = builtin_print._vadc_orig

    def print(*args, **kwargs):
        force = kwargs.pop('force', False)
        if is_master or force:
            builtin_print(*args, **kwargs)

    print._vadc_wrapped = True
    print._vadc_orig = builtin_print
    __builtin__.print = print


def init_distributed_mode(args):
    """utils/distritributed_model.py:38-70 with backend nccl (gloo when no GPU)."""
    if 'RANK' in os.environ and 'WORLD_SIZE' in os.environ:
        args.rank = int(os.environ["RANK"])
        args.world_size = int(os.environ['WORLD_SIZE'])
        args.gpu = int(os.environ.get('LOCAL_RANK', 0))
    elif 'SLURM_PROCID' in os.environ:
        args.rank = int(os.environ['SLURM_PROCID'])
        args.gpu = args.rank % max(torch.cuda.device_count(), 1)
        args.world_size = int(os.environ.get('SLURM_NTASKS', getattr(args, 'world_size', 1)))
    elif torch.cuda.is_available():
        print('Will run the code on one GPU.')
        args.rank, args.gpu, args.world_size = 0, 0, 1
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
    else:
        print('Does not support training without GPU.')
        sys.exit(1)

    use_cuda = torch.cuda.is_available()
    backend = "nccl" if use_cuda else "gloo"
    kwargs = {}
    if use_cuda:
        torch.cuda.set_device(args.gpu)
        kwargs["device_id"] = torch.device("cuda", args.gpu)
    dist.init_process_group(backend=backend, init_method=getattr(args, "dist_url", "env://"),
                            world_size=args.world_size, rank=args.rank, **kwargs)
    print('| distributed init (rank {}): {}'.format(args.rank, getattr(args, "dist_url", "env://")), flush=True)
    dist.barrier()
    setup_for_distributed(args.rank == 0)


def fix_random_seeds(seed=31):
    """utils/distritributed_model.py:73-79"""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)


def get_sha():
    """utils/distritributed_model.py:82-100"""
    cwd = os.path.dirname(os.path.abspath(__file__))

    def _run(command):
        return subprocess.check_output(command, cwd=cwd, stderr=subprocess.DEVNULL).decode('ascii').strip()

    sha, diff, branch = 'N/A', "clean", 'N/A'
    try:
        sha = _run(['git', 'rev-parse', 'HEAD'])
        diff = _run(['git', 'diff-index', 'HEAD'])
        diff = "has uncommited changes" if diff else "clean"
        branch = _run(['git', 'rev-parse', '--abbrev-ref', 'HEAD'])
    except Exception:
        pass
    return f"sha: {sha}, status: {diff}, branch: {branch}"


# ---------------------------------------------------------------------------
# hot-path collectives
# ---------------------------------------------------------------------------
def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(n_items, rank=None, world_size=None):
    """contiguous split of ``n_items`` clips over ranks (first ranks get the remainder)"""
    if rank is None:
        rank = dist.get_rank() if is_dist() else 0
    if world_size is None:
        world_size = dist.get_world_size() if is_dist() else 1
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class _PeerAllReduce:
    """One-shot all-reduce over NVLink peer memory (csrc/collective.cu): a symmetric block per rank — 256 flag words +
    2 x capacity floats — allocated and exchanged ONCE through torch.distributed._symmetric_memory (plumbing); after
    that every all-reduce of up to four small fp32 tensors is a single launch of our own kernel: no NCCL call, no packed
    staging tensor, capturable in a CUDA graph."""

    FLAG_WORDS = 256

    def __init__(self, capacity, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group or dist.group.WORLD
        self.capacity = int(capacity)
        self.device = torch.device("cuda", torch.cuda.current_device())
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        self.block = symm.empty(self.FLAG_WORDS + 2 * self.capacity, dtype=torch.float32, device=self.device)
        self.block.zero_()
        self.handle = symm.rendezvous(self.block, group)
        self.rank, self.world = self.handle.rank, self.handle.world_size
        self.peers_dev = int(self.handle.buffer_ptrs_dev)
        torch.cuda.synchronize()
        dist.barrier(group)                       # every block is zeroed before the first flag can arrive

    def all_reduce_(self, tensors):
        from . import _lib
        ts = [t for t in tensors if t.numel() > 0]
        padded = sum((t.numel() + 3) // 4 * 4 for t in ts)
        if len(ts) > 4 or padded > self.capacity or padded * 4 * self.world > 200 * 1024 or any(
                t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device for t in ts):
            return False
        args = []
        for i in range(4):
            if i < len(ts):
                args += [_lib.ptr(ts[i]), ts[i].numel()]
            else:
                args += [None, 0]
        import ctypes
        _lib.check(_lib.lib().vadc_oneshot_allreduce(ctypes.c_void_p(self.peers_dev), self.rank, self.world, self.capacity,
                                                     *args, _lib.stream()), "vadc_oneshot_allreduce")
        return True


_peer_ar = None
_peer_ar_error = None


def enable_peer_allreduce(capacity=1 << 16):
    """switch ``allreduce_sum_packed`` / ``global_frobenius`` from NCCL collectives to the one-shot NVLink kernel for
    messages of up to ``capacity`` floats (default 256 KB).  Collective: every rank calls it once after
    ``init_distributed_mode``.  Returns True when the peer path is active; on any failure (no CUDA, one rank, symmetric
    memory unavailable) NCCL stays in use and the reason is kept in ``peer_allreduce_status()``."""
    global _peer_ar, _peer_ar_error
    if not (is_dist() and torch.cuda.is_available()):
        _peer_ar_error = "not a multi-rank CUDA job"
        return False
    ok = torch.zeros(1, device="cuda")
    try:
        _peer_ar = _PeerAllReduce(capacity)
        ok += 1
    except Exception as e:                      # noqa: BLE001 - the NCCL path is the fallback for the plumbing, not for the math
        _peer_ar, _peer_ar_error = None, repr(e)[:300]
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # all ranks or none
    if float(ok) < 1:
        _peer_ar = None
        _peer_ar_error = _peer_ar_error or "a peer rank could not set up symmetric memory"
        return False
    _peer_ar_error = None
    return True


def disable_peer_allreduce():
    global _peer_ar
    _peer_ar = None


def peer_allreduce_status():
    return "one-shot NVLink kernel (vadc_oneshot_allreduce)" if _peer_ar is not None else f"NCCL ({_peer_ar_error or 'peer path not enabled'})"


def allreduce_sum_packed(tensors, average=False, async_op=False):
    """ONE all-reduce(sum) for a list of tensors: they are packed into a flat
    buffer, reduced, and copied back in place (centroid grads [K,C] + LN grads
    [2C] + loss scalars are latency-bound messages: one collective, not four).
    ``average=True`` divides by world size (DDP-compatible mean).  Returns the
    work handle when ``async_op`` (call ``finish()`` on it)."""
    if not is_dist():
        return _Done()
    if _peer_ar is not None and _peer_ar.all_reduce_(tensors):
        if average:
            for t in tensors:
                t.div_(dist.get_world_size())
        return _Done()
    flat = torch.cat([t.reshape(-1) for t in tensors])
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=async_op)

    def unpack():
        if average:
            flat.div_(dist.get_world_size())
        off = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n

    if async_op:
        return _Pending(work, unpack)
    unpack()
    return _Done()


class _Done:
    def finish(self):
        return None


class _Pending:
    def __init__(self, work, fn):
        self.work, self.fn = work, fn

    def finish(self):
        self.work.wait()
        self.fn()


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x; backward: identity (every rank's objective sees
    the global sum, SURVEY.md §8e 'global-batch semantics')."""

    @staticmethod
    def forward(ctx, x):
        y = x.clone()
        if is_dist() and not (_peer_ar is not None and _peer_ar.all_reduce_([y])):
            dist.all_reduce(y, op=dist.ReduceOp.SUM)
        return y

    @staticmethod
    def backward(ctx, g):
        return g


def global_frobenius(loss_sq, ddp_compat=False):
    """sqrt of the cross-rank sum of a local sum-of-squares (cluster loss
    backbone.py:94,98; pixel loss main_predict.py:273-275) so that N ranks
    reproduce the single-process full-batch value.  ``ddp_compat`` keeps the
    reference's DDP behaviour instead (per-rank norm)."""
    if ddp_compat or not is_dist():
        return torch.sqrt(loss_sq)
    return torch.sqrt(_AllReduceSum.apply(loss_sq))
