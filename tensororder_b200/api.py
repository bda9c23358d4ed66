"""`B200API` — the `--tensor_library=b200` backend.

Mirrors the reference's tensor-library interface for the execution path
(`BaseTensorAPI`, src/tensor_network/tensor_apis/base_api.py:8-28, and the de-facto members of
`NumpyAPI`, src/tensor_network/tensor_apis/numpy_apis.py:24-65): same method names, argument
meaning and error behaviour, so `execution.run` (src/execution.py:122-152) drives it unchanged.
All arithmetic happens in the CUDA library behind include/tob200.h; there is no CPU fallback.

Multi-GPU: when `torch.distributed` is initialised, rank r contracts slices
r, r+W, r+2W, ... and the partial counts are combined with one all-reduce (NCCL on GPUs)."""
from __future__ import annotations

import ctypes
import time
from ctypes import byref, c_double, c_float, c_int32, c_void_p
from typing import Optional

import numpy as np

from . import cabi
from .flatten import FlatPlan, flatten_plan


class OutOfMemoryError(Exception):
    """Stands in for `tensor_network.OutOfMemoryError` (base_api.py:4) when the reference is not
    importable; `B200API` raises the reference's own class when it is."""


def _oom_class():
    try:  # the reference package, when this backend is registered inside TensorOrder
        from tensor_network.tensor_apis.base_api import OutOfMemoryError as RefOOM  # type: ignore

        return RefOOM
    except Exception:
        return OutOfMemoryError


def _ptr(arr: np.ndarray, ctype):
    return arr.ctypes.data_as(ctypes.POINTER(ctype))


def _primes_below(limit: int, count: int):
    """The `count` largest primes below `limit` (trial division; limit is 2^23, so divisors < 2897)."""
    out = []
    n = limit - 1
    while len(out) < count:
        if n % 2 and all(n % d for d in range(3, int(n ** 0.5) + 1, 2)):
            out.append(n)
        n -= 1
    return out


EXACT_PRIMES = _primes_below(1 << 23, 96)  # ~23 bits each: enough for counts up to 2^2200


def crt(residues, moduli) -> int:
    """Chinese remainder: the integer in [0, prod(moduli)) with the given residues."""
    x, m = 0, 1
    for r, p in zip(residues, moduli):
        t = ((r - x) * pow(m, -1, p)) % p
        x += m * t
        m *= p
    return x


class CompiledPlan:
    """Owns one `tob_plan` (device arena, leaf tensors, CUDA graph)."""

    def __init__(self, flat: FlatPlan, device: int = 0, use_graph=None, kernel_policy: int = 0,
                 hoist_invariant: bool = True, mem_limit_bytes: int = 0, use_microtree: bool = True,
                 slice_lanes: int = 0, dag_branches: int = 0):
        self.flat = flat
        self._handle = c_void_p()
        desc = cabi.tob_plan_desc()
        desc.n_nodes = flat.n_nodes
        desc.node_left = _ptr(flat.node_left, c_int32)
        desc.node_right = _ptr(flat.node_right, c_int32)
        desc.node_leaf = _ptr(flat.node_leaf, c_int32)
        desc.n_leaves = flat.n_leaves
        desc.leaf_rank = _ptr(flat.leaf_rank, c_int32)
        desc.leaf_data_offset = _ptr(flat.leaf_data_offset, ctypes.c_int64)
        desc.leaf_axis_start = _ptr(flat.leaf_axis_start, c_int32)
        desc.leaf_axis_edge = _ptr(flat.leaf_axis_edge, c_int32)
        desc.n_slice_groups = flat.n_slice_groups
        desc.leaf_data_len = int(flat.leaf_data.shape[0])
        opt = cabi.tob_options()
        cabi.lib.tob_default_options(byref(opt))
        opt.device = device
        if use_graph is not None:  # None keeps the library default (auto)
            opt.use_graph = int(use_graph) if not isinstance(use_graph, bool) else (1 if use_graph else 0)
        opt.kernel_policy = kernel_policy
        opt.hoist_invariant = 1 if hoist_invariant else 0
        opt.mem_limit_bytes = int(mem_limit_bytes)
        opt.use_microtree = 1 if use_microtree else 0
        opt.slice_lanes = int(slice_lanes)
        opt.dag_branches = int(dag_branches)
        rc = cabi.lib.tob_plan_create(byref(desc), byref(opt), byref(self._handle))
        if rc != cabi.TOB_OK:
            raise ValueError("tob_plan_create: " + cabi.last_error())
        self.uploaded = False

    # -- host-only queries --
    @property
    def num_slices(self) -> int:
        return int(cabi.lib.tob_plan_num_slices(self._handle))

    @property
    def peak_bytes(self) -> int:
        return int(cabi.lib.tob_plan_peak_bytes(self._handle))

    @property
    def num_ops(self) -> int:
        return int(cabi.lib.tob_plan_num_ops(self._handle))

    def describe(self) -> dict:
        import json

        n = cabi.lib.tob_plan_describe(self._handle, None, 0)
        buf = ctypes.create_string_buffer(n + 1)
        cabi.lib.tob_plan_describe(self._handle, buf, n + 1)
        return json.loads(buf.value.decode())

    # -- device --
    def upload(self) -> None:
        rc = cabi.lib.tob_plan_upload(self._handle, _ptr(self.flat.leaf_data, c_double), int(self.flat.leaf_data.shape[0]))
        if rc == cabi.TOB_E_OOM:
            raise _oom_class()(cabi.last_error())
        if rc != cabi.TOB_OK:
            raise RuntimeError("tob_plan_upload: " + cabi.last_error())
        self.uploaded = True

    def run(self, first: int = 0, count: Optional[int] = None, stride: int = 1, initial: float = 0.0,
            skip_invariant: bool = False) -> float:
        if count is None:
            count = (self.num_slices - first + stride - 1) // stride
        out = c_double(0.0)
        rc = cabi.lib.tob_plan_run_ex(self._handle, first, count, stride, float(initial), 1 if skip_invariant else 0,
                                      byref(out))
        if rc == cabi.TOB_E_OOM:
            raise _oom_class()(cabi.last_error())
        if rc != cabi.TOB_OK:
            raise RuntimeError("tob_plan_run: " + cabi.last_error())
        return out.value

    def run_interruptible(self, first: int = 0, count: Optional[int] = None, stride: int = 1, target_s: float = 0.25):
        """Same result as `run` (bit for bit: the device accumulator is carried across calls), but
        returns to the interpreter every ~target_s seconds of device work so pending signal handlers —
        the reference's SIGALRM `TimeoutTimer` (src/util/util.py:32-39) — run between chunks of slices."""
        if count is None:
            count = (self.num_slices - first + stride - 1) // stride
        acc, done, chunk = 0.0, 0, 1
        total_ms, launches = 0.0, 0
        gemm = [0.0, 0.0, 0]
        while done < count or (count == 0 and done == 0):
            c = min(chunk, count - done)
            acc = self.run(first + done * stride, c, stride, initial=acc, skip_invariant=done > 0)
            ms = self.last_ms
            total_ms += ms
            launches += self.last_launches
            g = self.last_gemm
            gemm = [gemm[0] + g[0], gemm[1] + g[1], gemm[2] + g[2]]
            done += max(c, 1)
            if c > 0 and ms > 0:
                chunk = max(1, min(count, int(c * target_s * 1e3 / ms)))
            if count == 0:
                break
        self.interruptible_stats = {"device_ms": total_ms, "launches": launches, "gemm": tuple(gemm)}
        return acc

    @property
    def last_ms(self) -> float:
        return float(cabi.lib.tob_plan_last_ms(self._handle))

    @property
    def last_issue_ms(self) -> float:
        return float(cabi.lib.tob_plan_last_issue_ms(self._handle))

    @property
    def last_launches(self) -> int:
        return int(cabi.lib.tob_plan_last_launches(self._handle))

    @property
    def last_gemm(self):
        """(ms, flops, launches) of the DMMA GEMM kernels in the last run (stream mode only)."""
        ms, fl, n = c_double(0), c_double(0), ctypes.c_int64(0)
        cabi.lib.tob_plan_last_gemm(self._handle, byref(ms), byref(fl), byref(n))
        return ms.value, fl.value, n.value

    def set_modulus(self, modulus: int) -> None:
        rc = cabi.lib.tob_plan_set_modulus(self._handle, float(modulus))
        if rc != cabi.TOB_OK:
            raise ValueError("tob_plan_set_modulus: " + cabi.last_error())

    def set_gemm_timing(self, on: bool = True) -> None:
        cabi.lib.tob_plan_set_gemm_timing(self._handle, 1 if on else 0)

    def set_stream(self, cuda_stream_handle: int) -> None:
        rc = cabi.lib.tob_plan_set_stream(self._handle, c_void_p(cuda_stream_handle))
        if rc != cabi.TOB_OK:
            raise RuntimeError("tob_plan_set_stream: " + cabi.last_error())

    def profile(self, slice_id: int = 0):
        n = self.num_ops
        ms = (c_float * n)()
        out = c_double(0.0)
        rc = cabi.lib.tob_plan_profile(self._handle, slice_id, ms, n, byref(out))
        if rc != cabi.TOB_OK:
            raise RuntimeError("tob_plan_profile: " + cabi.last_error())
        return list(ms), out.value

    def close(self) -> None:
        if self._handle:
            cabi.lib.tob_plan_destroy(self._handle)
            self._handle = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class B200API:
    """Drop-in for `tensor_network.ALL_APIS[...]` entries (src/tensor_network/__init__.py:12-16)."""

    def __init__(self):
        self._entry_type = "float64"
        self._device = None  # None: LOCAL_RANK under torchrun, else 0
        self._use_graph = None  # library default: graph replay for launch-bound slices
        self._kernel_policy = 0
        self._hoist = True
        self._microtree = True
        self._lanes = 0
        self._branches = 0
        self._distributed = True
        self.last_stats = {}

    # ---- configuration (tensororder.py:205-217, execution.py:82-85) ----
    def add_argument(self, key, value):
        if key == "entry_type":
            # float64: the DMMA path.  bigint: the reference's exact mode (numpy object arrays of Python ints,
            # numpy_apis.py:21) done as residues modulo ~23-bit primes on the same kernels + CRT on the host.
            if value not in ("float64", "bigint"):
                raise ValueError("Unknown b200 type %s (float64 and bigint are implemented)" % value)
            self._entry_type = value
        elif key == "thread_limit":
            pass  # BLAS thread cap of the numpy backend; no host threads are used here
        elif key == "device":
            self._device = int(value)
        elif key == "use_graph":
            self._use_graph = value
        elif key == "kernel_policy":
            self._kernel_policy = int(value)
        elif key == "hoist_invariant":
            self._hoist = bool(value)
        elif key == "slice_lanes":
            self._lanes = int(value)
        elif key == "dag_branches":
            self._branches = int(value)
        elif key == "use_microtree":
            self._microtree = bool(value)
        elif key == "distributed":
            self._distributed = bool(value)
        else:
            # same message and exception type as BaseTensorAPI.add_argument (base_api.py:9-12)
            raise ValueError("Invalid argument " + str(key) + " for selected tensor_library")

    def get_entry_size(self):
        return 8  # numpy.dtype(float64).itemsize, numpy_apis.py:64-65

    def warm(self):
        self._resolve_device()
        if cabi.lib.tob_device_count() <= 0:
            raise RuntimeError("b200 tensor library: no CUDA device")

    # ---- leaf construction: host arrays, like every reference backend (numpy_apis.py:36-40) ----
    def create_tensor(self, shape, default_value=None):
        if default_value is None:
            return np.empty(shape, dtype=np.float64)
        return np.full(shape, default_value, dtype=np.float64)

    # ---- the primary entry (base_api.py:17-28, called from execution.py:136) ----
    def contract_sliced(self, execution_plan, num_slice_limit=None):
        if self._entry_type == "bigint":
            return self._contract_exact(execution_plan, num_slice_limit)
        t0 = time.perf_counter()
        flat = flatten_plan(execution_plan)  # leaves are built through a buffer-backed create_tensor
        rank, world = self._rank_world()
        compiled = CompiledPlan(flat, device=self._resolve_device(), use_graph=self._use_graph,
                                kernel_policy=self._kernel_policy, hoist_invariant=self._hoist,
                                use_microtree=self._microtree, slice_lanes=self._lanes,
                                dag_branches=self._branches)
        try:
            t1 = time.perf_counter()
            compiled.upload()
            t2 = time.perf_counter()
            total = compiled.num_slices
            if num_slice_limit is not None:
                total = min(total, int(num_slice_limit))  # itertools.islice(slices, N), base_api.py:23-24
            count = 0 if rank >= total else (total - rank + world - 1) // world
            partial = compiled.run_interruptible(first=rank if count else 0, count=count, stride=world)
            t3 = time.perf_counter()
            result = self._all_reduce(partial) if world > 1 else partial
            self.last_stats = {
                "flatten_compile_s": t1 - t0, "upload_s": t2 - t1, "run_s": t3 - t2,
                "device_ms": compiled.interruptible_stats["device_ms"],
                "launches": compiled.interruptible_stats["launches"],
                "h2d_bytes": int(flat.leaf_data.nbytes), "d2h_bytes": 32,
                "peak_bytes": compiled.peak_bytes, "slices": total, "rank": rank, "world": world,
            }
        finally:
            compiled.close()
        return np.float64(result)

    # ---- exact counts (entry_type = bigint) ----
    def _contract_exact(self, execution_plan, num_slice_limit=None):
        """Exact integer result, any magnitude: one float64 pass sizes the answer, then the same compiled
        plan runs once per prime p < 2^23 with every kernel reducing modulo p (all sums stay below 2^53, so
        FP64 arithmetic is exact), and the residues are combined by CRT.  Primes are added until the
        reconstruction is stable and agrees with the float64 estimate."""
        import dataclasses
        import math

        t0 = time.perf_counter()
        flat = flatten_plan(execution_plan)
        if not np.all(np.isfinite(flat.leaf_data)) or not np.array_equal(flat.leaf_data, np.rint(flat.leaf_data)) \
                or np.any(np.abs(flat.leaf_data) >= 2.0 ** 53):
            raise ValueError("entry_type bigint needs integer-valued tensors (weights must be integers)")
        rank, world = self._rank_world()
        compiled = CompiledPlan(flat, device=self._resolve_device(), use_graph=self._use_graph,
                                kernel_policy=self._kernel_policy, hoist_invariant=self._hoist,
                                use_microtree=self._microtree, slice_lanes=self._lanes,
                                dag_branches=self._branches)
        try:
            total = compiled.num_slices
            if num_slice_limit is not None:
                total = min(total, int(num_slice_limit))
            count = 0 if rank >= total else (total - rank + world - 1) // world
            first = rank if count else 0

            def one_pass(leaf_data, modulus):
                compiled.flat = dataclasses.replace(flat, leaf_data=np.ascontiguousarray(leaf_data))
                compiled.upload()
                compiled.set_modulus(modulus)
                partial = compiled.run_interruptible(first=first, count=count, stride=world)
                return self._all_reduce(partial) if world > 1 else partial

            estimate = float(one_pass(flat.leaf_data, 0))
            if math.isfinite(estimate):
                bits = math.log2(abs(estimate) + 1.0) + 4.0
            else:  # beyond float64: a-priori bound, the product over tensors of their 1-norms
                bits = 4.0
                for l in range(flat.n_leaves):
                    off = int(flat.leaf_data_offset[l])
                    bits += math.log2(max(float(np.abs(flat.leaf_data[off: off + (1 << int(flat.leaf_rank[l]))]).sum()), 1.0))
            residues, moduli = [], []
            value, passes = None, 0
            need = max(1, math.ceil(bits / math.log2(EXACT_PRIMES[-1])))
            while True:
                if len(moduli) >= len(EXACT_PRIMES):
                    raise OverflowError("count needs more than %d primes" % len(EXACT_PRIMES))
                p = EXACT_PRIMES[len(moduli)]
                r = one_pass(np.mod(flat.leaf_data, p), p)
                passes += 1
                if r != int(r) or not (0 <= r < p * max(world, 1)):
                    raise RuntimeError("exact mode: residue %r is not an integer below the modulus" % r)
                residues.append(int(r) % p)
                moduli.append(p)
                if len(moduli) < need:
                    continue
                x = crt(residues, moduli)
                m = math.prod(moduli)
                if x > m // 2 and (not math.isfinite(estimate) or estimate < 0):
                    x -= m  # symmetric range for negative totals
                # stable (the last prime did not change the reconstruction) and consistent with the float64 pass
                prev = crt(residues[:-1], moduli[:-1]) if len(moduli) > 1 else None
                if prev is not None and prev > math.prod(moduli[:-1]) // 2 and x < 0:
                    prev -= math.prod(moduli[:-1])
                close = (not math.isfinite(estimate)) or abs(x - estimate) <= 1e-6 * abs(estimate) + 1.0
                if close and (prev == x or len(moduli) == 1 and abs(x) < m // 4):
                    value = x
                    break
            self.last_stats = {"exact_passes": passes, "primes": list(moduli), "float_estimate": estimate,
                               "seconds": time.perf_counter() - t0, "slices": total, "rank": rank, "world": world,
                               "launches": compiled.interruptible_stats["launches"]}
        finally:
            compiled.close()
        return value

    # ---- secondary entries ----
    def contract(self, network, contraction_tree):
        """One (unsliced) network; returns a 0-d array so `result[tuple()]` works (base_api.py:26-27)."""

        class _Plan:
            pass

        plan = _Plan()
        plan.tree, plan.network, plan.groups_to_slice = contraction_tree, network, []
        keep = self._distributed
        self._distributed = False
        try:
            value = self.contract_sliced(plan)
            # 0-d array so `result[tuple()]` works; exact mode keeps the Python int (object dtype)
            return np.array(value, dtype=object if self._entry_type == "bigint" else np.float64)
        finally:
            self._distributed = keep

    def tensordot(self, a, b, axes):
        """`numpy.tensordot(a, b, (axes_a, axes_b))` on the GPU for host arrays whose axes all have
        extent 2 (call site tensor_network.pyx:152-154)."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        axes_a, axes_b = axes
        axes_a = [int(x) % max(a.ndim, 1) for x in np.atleast_1d(axes_a)] if np.size(axes_a) else []
        axes_b = [int(x) % max(b.ndim, 1) for x in np.atleast_1d(axes_b)] if np.size(axes_b) else []
        if len(axes_a) != len(axes_b):
            raise ValueError("shape-mismatch for sum")
        if any(s != 2 for s in a.shape) or any(s != 2 for s in b.shape):
            raise ValueError("b200 tensordot: every index must have extent 2")
        rank_c = a.ndim + b.ndim - 2 * len(axes_a)
        c = np.empty((2,) * rank_c, dtype=np.float64)
        aa = np.asarray(axes_a, dtype=np.int32)
        ab = np.asarray(axes_b, dtype=np.int32)
        rc = cabi.lib.tob_tensordot_host(_ptr(a, c_double), a.ndim, _ptr(b, c_double), b.ndim, _ptr(aa, c_int32),
                                         _ptr(ab, c_int32), len(axes_a), _ptr(c, c_double))
        if rc == cabi.TOB_E_OOM:
            raise _oom_class()(cabi.last_error())
        if rc != cabi.TOB_OK:
            raise RuntimeError("tob_tensordot_host: " + cabi.last_error())
        return c

    # ---- helpers ----
    def _resolve_device(self) -> int:
        if self._device is not None:
            return self._device
        import os

        return int(os.environ.get("LOCAL_RANK", "0")) if self._dist_ready() else 0

    def _dist_ready(self) -> bool:
        if not self._distributed:
            return False
        try:
            import torch.distributed as dist

            return dist.is_available() and dist.is_initialized()
        except Exception:
            return False

    def _rank_world(self):
        if self._dist_ready():
            import torch.distributed as dist

            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _all_reduce(self, partial: float) -> float:
        """One all-reduce of a float64 scalar (the only exchange step of the sliced path)."""
        import torch
        import torch.distributed as dist

        backend = dist.get_backend()
        dev = torch.device("cuda", self._resolve_device()) if backend == "nccl" else torch.device("cpu")
        t = torch.tensor([partial], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())


B200_APIS = {"b200": B200API}


def register(all_apis=None):
    """Adds "b200" to the reference's live registry `tensor_network.ALL_APIS`
    (src/tensor_network/__init__.py:12-16) — the dict `util.TaggedChoice` resolves `--tensor_library`
    against (src/tensororder.py:88-94, src/util/util.py:249-252)."""
    if all_apis is None:
        import tensor_network  # type: ignore  (the reference package; needs <TensorOrder>/src on sys.path)

        all_apis = tensor_network.ALL_APIS
    all_apis.update(B200_APIS)
    return all_apis
