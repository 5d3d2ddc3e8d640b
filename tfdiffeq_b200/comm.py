"""Shared-step groups: independent batch shards on several GPUs taking ONE step sequence.

The reference has no distributed code; it folds every batch axis into one ODE system with a single step
size and a tolerance that is a *global scalar* over the whole tensor (tfdiffeq/misc.py:257).  To keep that
semantics when the batch is sharded across GPUs, every attempt needs, per tuple component,
``{sum err^2, max|y0|, max|y1|, non-finite}`` over all ranks -- 32 bytes per rank.  ``libb2ode`` exchanges
them inside the finalize kernel itself: the last block of each rank stores its totals into every peer's
mailbox over NVLink (peer-mapped through CUDA IPC), spins on the arrival flags, and combines in rank
order, so all ranks take bit-identical accept / dt decisions with no host involvement and no NCCL call on
the hot path.  ``torch.distributed`` (NCCL or gloo) is used once, at setup, to exchange the IPC handles.

Usage (one process per GPU)::

    group = SharedStepGroup()                       # uses the default process group
    y_local = odeint(func, y0_local, t, method='dopri5', options={'shared_step_group': group})

``shard_batch`` gives the contiguous split of the leading batch axis the benchmarks use.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def bind_to_gpu_numa(device=None):
    """Pin the calling process (and therefore the page-locked buffers it allocates from now on: first touch) to the NUMA
    node the GPU hangs off.  On an 8-GPU box GPUs 4-7 sit on the second socket; a rank whose pinned staging buffers live on
    the other socket pushes every host<->device byte across the inter-socket link (round 1: end-to-end time per solve went
    from 34 ms at 1-4 GPUs to 108 ms at 8).  Returns the node number, or None when the topology cannot be read (no
    /sys entry, container without the PCI tree): the caller carries on unbound."""
    import os
    try:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        props = torch.cuda.get_device_properties(dev)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                if "-" in part:
                    lo, hi = part.split("-")
                    cpus.update(range(int(lo), int(hi) + 1))
                elif part:
                    cpus.add(int(part))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def shard_bounds(n, world_size, rank):
    """Contiguous, balanced split of ``n`` items: the first ``n % world_size`` ranks get one extra."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(y, world_size, rank):
    lo, hi = shard_bounds(y.shape[0], world_size, rank)
    return y[lo:hi]


def combine_partials(per_rank):
    """Host-side statement of what the kernel's ``group_combine`` computes (rank order, NaN-propagating max);
    used by the CPU tests to check the sharded error norm against the unsharded one.

    per_rank: list over ranks of (sum_sq, max0, max1, bad) tuples."""
    tot = [0.0, 0.0, 0.0, 0.0]
    for q, (s, m0, m1, bad) in enumerate(per_rank):
        if q == 0:
            tot = [s, m0, m1, bad]
            continue
        tot[0] = tot[0] + s
        for c, v in ((1, m0), (2, m1), (3, bad)):
            tot[c] = float("nan") if (tot[c] != tot[c] or v != v) else max(tot[c], v)
    return tuple(tot)


class SharedStepGroup(object):
    """Owns this rank's mailbox and the peer mappings; attach()es them to each native solver."""

    def __init__(self, group=None, device=None):
        if not dist.is_initialized():
            raise RuntimeError("SharedStepGroup needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > _lib.MAXPEERS:
            raise ValueError("a shared-step group spans at most %d GPUs (one NVSwitch box)" % _lib.MAXPEERS)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._own = C.c_void_p()
        self._peers = []
        self._sums = {}
        lib, check = _lib.lib, _lib.check
        with torch.cuda.device(self.device):
            handle = C.create_string_buffer(64)
            check(lib.b2ode_mailbox_create(C.byref(self._own), handle))
            mine = torch.tensor(list(handle.raw), dtype=torch.uint8)
            backend = dist.get_backend(group)
            if backend == "nccl":
                mine = mine.to(self.device)
            gathered = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine, group=group)
            self._ptrs = _lib.PtrArray()
            for r in range(self.world):
                if r == self.rank:
                    self._ptrs[r] = self._own.value
                    self._peers.append(None)
                else:
                    p = C.c_void_p()
                    check(lib.b2ode_mailbox_open(bytes(gathered[r].cpu().tolist()), C.byref(p)))
                    self._ptrs[r] = p.value
                    self._peers.append(p)
            dist.barrier(group=group)

    # Group-wide sums of per-rank integers are needed once per *shape*, not once per solve: the collective (and its
    # host synchronisation) runs the first time a rank presents a given local tuple and is cached afterwards.
    # Contract (SPMD): every rank of the group calls odeint with the same sequence of problems, so either all ranks hit
    # their cache or all ranks miss it.  The key is the LOCAL shape: a sequence in which only some ranks' local shapes
    # change (101 rows over two ranks = 51 + 50, then 102 = 51 + 51: rank 0 sees 51 twice) breaks the contract -- shard such
    # batches so that every rank's size changes, or use a fresh group.
    def _sum_cached(self, tag, values):
        key = (tag,) + tuple(int(v) for v in values)
        hit = self._sums.get(key)
        if hit is None:
            v = torch.tensor([int(x) for x in values], dtype=torch.int64)
            if dist.get_backend(self.group) == "nccl":
                v = v.to(self.device)
            dist.all_reduce(v, group=self.group)
            hit = self._sums[key] = tuple(int(x) for x in v.cpu().tolist())
        return hit

    def attach(self, handle, seg, replicated=()):
        """Called by the solver after ``b2ode_adaptive_bind``.  ``replicated``: indices of tuple components every rank
        holds in full (identical values on all ranks, e.g. the parameter adjoint after its all-reduce): they are
        counted once in the error norm instead of being summed over the ranks."""
        lib, check = _lib.lib, _lib.check
        # group-wide element count per component: the mean in misc.py:262 runs over every rank's elements
        tot = self._sum_cached("lens", seg.lens)
        glob = _lib.LenArray(*[(seg.lens[i] if i in replicated else tot[i]) for i in range(len(seg.lens))])
        check(lib.b2ode_comm_attach(handle, self.rank, self.world, self._ptrs))
        check(lib.b2ode_comm_set_global_len(handle, glob))
        mask = 0
        for i in replicated:
            mask |= 1 << i
        check(lib.b2ode_comm_set_replicated(handle, mask))

    def global_count(self, n):
        """Sum of a per-rank count over the group (e.g. trajectories, for the mean in the error norm)."""
        return self._sum_cached("count", (n,))[0]

    def agree_fused(self, n_traj, fits):
        """(trajectories of every rank, every shard fits the persistent fused kernel): all shards must take the same path
        -- the fused kernel and the generic kernels use different exchange protocols -- and each rank derives every
        other rank's kernel grid from its shard size.  One all-gather per distinct local (n_traj, fits), cached."""
        key = ("fused", int(n_traj), bool(fits))
        hit = self._sums.get(key)
        if hit is None:
            v = torch.tensor([int(n_traj), 1 if fits else 0], dtype=torch.int64)
            if dist.get_backend(self.group) == "nccl":
                v = v.to(self.device)
            out = [torch.empty_like(v) for _ in range(self.world)]
            dist.all_gather(out, v, group=self.group)
            rows = [o.cpu().tolist() for o in out]
            hit = self._sums[key] = ([int(r[0]) for r in rows], all(int(r[1]) == 1 for r in rows))
        return hit

    def close(self):
        lib = _lib.lib
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for p in self._peers:
                if p is not None:
                    lib.b2ode_mailbox_close(p)
            self._peers = []
            if self._own:
                lib.b2ode_mailbox_destroy(self._own)
                self._own = C.c_void_p()
