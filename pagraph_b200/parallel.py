"""Data-parallel plumbing of the trainer (reference: examples/profile/pa_gcn.py:18-24,65,86-97).

The reference wraps the 23 k-parameter model in DistributedDataParallel; the only cross-GPU traffic
is the ~92 KB gradient all-reduce. Here that is ONE flat fp32 bucket and ONE all-reduce per step on
the compute stream (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests): at this size the
collective is pure latency, so DDP's per-bucket hooks buy nothing. Partitions are self-sufficient
(PaGraph/partition/utils.py:9-52), so no feature or activation ever crosses GPUs.

Also here: the hash split of train ids (PaGraph/partition/hash.py:25-51) and the per-rank batch
count equalisation the reference lacks (an uneven count hangs its all-reduce).
"""
import numpy as np
import torch
import torch.distributed as dist


class FlatGradAllReduce:
    """Averages the gradients of `module` across ranks with a single all-reduce of one flat buffer.

    Parameters are re-pointed at views of one flat tensor (and so are their .grad), so the
    all-reduce needs no pack/unpack kernels. Usage, in place of DDP:
        sync = FlatGradAllReduce(model); ...; loss.backward(); sync(); optimizer.step()
    """

    def __init__(self, module, process_group=None):
        self.group = process_group
        self.params = [p for p in module.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat_param = torch.empty(n, dtype=ref.dtype, device=ref.device)
        self.flat_grad = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_param[off:off + k].view_as(p)
                p.grad = self.flat_grad[off:off + k].view_as(p)
                off += k
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        if self.world > 1:   # same initial weights everywhere (DDP broadcasts rank 0's at construction)
            dist.broadcast(self.flat_param, src=0, group=process_group)

    def flat_parameters(self):
        """[one nn.Parameter] aliasing every parameter (and its .grad aliasing every gradient): hand it to an
        elementwise optimizer (SGD / Adam / AdamW) instead of module.parameters() and the update is a single kernel over
        the flat bucket; the result is identical because those optimizers act elementwise."""
        if getattr(self, "_flat", None) is None:
            self._flat = torch.nn.Parameter(self.flat_param, requires_grad=True)
            self._flat.grad = self.flat_grad
        return [self._flat]

    def zero_grad(self):
        """Keeps .grad pointing into the flat bucket (optimizer.zero_grad(set_to_none=True) would not)."""
        self.flat_grad.zero_()

    def __call__(self):
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            self.flat_grad.div_(self.world)


def hash_split(train_nids, num_parts, seed=None):
    """hash.py:25-51: shuffle the train ids, cut them into `num_parts` equal consecutive chunks (the
    last one takes the remainder). `seed` makes it reproducible (the reference shuffles unseeded)."""
    ids = np.array(train_nids, dtype=np.int64, copy=True)
    rng = np.random.default_rng(seed) if seed is not None else np.random
    rng.shuffle(ids)
    size = len(ids) // num_parts
    parts = []
    for p in range(num_parts):
        lo = p * size
        hi = (p + 1) * size if p != num_parts - 1 else len(ids)
        parts.append(ids[lo:hi])
    return parts


def equalised_num_batches(local_num_batches, process_group=None):
    """min over ranks of the per-epoch batch count, so that every rank enters the same number of
    all-reduces (the reference's remote-sampling path pads instead, parallel/dataloader.py:138-143)."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return int(local_num_batches)
    dev = "cuda" if dist.get_backend(process_group) == "nccl" else "cpu"
    t = torch.tensor([int(local_num_batches)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=process_group)
    return int(t.item())


class PeerAdam:
    """Gradient all-reduce + Adam step as one kernel over NVLink peer memory (pg_allreduce_adam), standing where
    `sync(); optimizer.step()` stands. The optimizer object keeps owning the hyper-parameters and the state tensors
    (`exp_avg`, `exp_avg_sq`, `step`), which the kernel updates in place, so `optimizer.state_dict()` stays meaningful.

    Usable when `optimizer` is a torch.optim.Adam (no amsgrad / maximize) over `sync.flat_parameters()` on a CUDA device,
    with world_size <= 8 on one node. Hyper-parameters are read when `step()` is called (a captured CUDA graph bakes them)."""

    @staticmethod
    def supported(sync, optimizer):
        if not isinstance(optimizer, torch.optim.Adam) or len(optimizer.param_groups) != 1:
            return False
        g = optimizer.param_groups[0]
        flat = sync.flat_parameters()[0]
        return (len(g["params"]) == 1 and g["params"][0] is flat and flat.is_cuda and flat.dtype == torch.float32
                and not g.get("amsgrad") and not g.get("maximize") and not g.get("decoupled_weight_decay", False)
                and sync.world <= 8)

    def __init__(self, sync, optimizer):
        import ctypes
        from . import _lib
        self._lib = _lib
        self.sync, self.opt = sync, optimizer
        self.flat = sync.flat_parameters()[0]
        dev = self.flat.device
        st = optimizer.state[self.flat]
        if len(st) == 0:                                  # what Adam._init_group creates for capturable / fused
            st["step"] = torch.zeros((), dtype=torch.float32, device=dev)
            st["exp_avg"] = torch.zeros_like(self.flat, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(self.flat, memory_format=torch.preserve_format)
        if not (torch.is_tensor(st["step"]) and st["step"].is_cuda):
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32, device=dev)
        self.state = st
        world, rank = sync.world, (dist.get_rank(sync.group) if sync.world > 1 else 0)
        h = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * _lib.PG_IPC_HANDLE_BYTES)()
        self._handle = None
        # Every rank walks the same sequence of collectives whatever happens locally, and the outcome is agreed with an
        # all-reduce(MIN) of a success flag: if CUDA IPC / peer access fails on ONE rank, all ranks raise (and fall back
        # to NCCL together) instead of some of them spinning in the kernel on flags that never arrive.
        err = None
        try:
            _lib.check(_lib.lib().pg_peer_group_create(world, rank, self.flat.numel(), dev.index, ctypes.byref(h), handle),
                       "pg_peer_group_create")
            self._handle = h
        except Exception as e:
            err = e
        if world > 1:
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine, group=sync.group)
            if err is None:
                try:
                    blob = bytes(torch.stack(allh).cpu().numpy().tobytes())
                    _lib.check(_lib.lib().pg_peer_group_connect(h, blob), "pg_peer_group_connect")
                except Exception as e:
                    err = e
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=sync.group)
            if int(ok.item()) == 0:
                self.close()
                raise RuntimeError("peer-memory all-reduce unavailable on at least one rank%s"
                                   % ("" if err is None else " (this rank: %s)" % err))
        elif err is not None:
            raise err

    def step(self, step_id, advance=False):
        """step_id: int64 CUDA scalar tensor, equal on all ranks. advance=False: the caller has already incremented it for
        this step (>= 1, +1 per call). advance=True: it holds the number of steps taken so far and the kernel increments
        it — and the optimizer's own `step` — itself (pg_allreduce_adam_next: two one-element kernels fewer per step)."""
        _lib = self._lib
        g = self.opt.param_groups[0]
        if not advance:
            self.state["step"].add_(1)
        fn, name = ((_lib.lib().pg_allreduce_adam_next, "pg_allreduce_adam_next") if advance else
                    (_lib.lib().pg_allreduce_adam, "pg_allreduce_adam"))
        with torch.cuda.device(self.flat.device):
            _lib.check(fn(self._handle, _lib.ptr(self.flat), _lib.ptr(self.sync.flat_grad), _lib.ptr(self.state["exp_avg"]),
                          _lib.ptr(self.state["exp_avg_sq"]), _lib.ptr(self.state["step"]), _lib.ptr(step_id), float(g["lr"]),
                          float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                          _lib.stream_ptr()), name)

    def close(self):
        if self._handle is not None:
            self._lib.lib().pg_peer_group_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
