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
import dataclasses
import threading
import time
from collections import OrderedDict
from ctypes import byref, c_double, c_float, c_int32, c_void_p
from typing import Optional

import numpy as np

from . import cabi
from .flatten import FlatPlan, flatten_plan, rebuild_leaf_data


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
        self.busy = 0  # plan cache: a call is running on this plan (another host thread must not share or evict it)
        self.interruptible_stats = {"device_ms": 0.0, "launches": 0, "gemm": (0.0, 0.0, 0)}
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

    def update_leaves(self, leaf_data: Optional[np.ndarray] = None) -> None:
        """Re-reads the leaf values into an uploaded plan (arena, tables and captured graphs are kept);
        a full `upload` when the plan holds no device memory."""
        if leaf_data is not None:
            self.flat = dataclasses.replace(self.flat, leaf_data=np.ascontiguousarray(leaf_data, dtype=np.float64))
        rc = cabi.lib.tob_plan_update_leaves(self._handle, _ptr(self.flat.leaf_data, c_double), int(self.flat.leaf_data.shape[0]))
        if rc == cabi.TOB_E_OOM:
            raise _oom_class()(cabi.last_error())
        if rc != cabi.TOB_OK:
            raise RuntimeError("tob_plan_update_leaves: " + cabi.last_error())
        self.uploaded = True

    def release_device(self) -> None:
        """Gives the device arena back (to the library's block pool) and keeps the compiled program."""
        cabi.lib.tob_plan_release(self._handle)
        self.uploaded = False

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

    def run_async(self, first: int = 0, count: Optional[int] = None, stride: int = 1, after_stream: int = 0) -> None:
        """Issues the run and returns (several plans can be in flight on one GPU); `wait` collects the result.
        after_stream: a cudaStream_t handle the run is ordered behind."""
        if count is None:
            count = (self.num_slices - first + stride - 1) // stride
        rc = cabi.lib.tob_plan_run_async(self._handle, first, count, stride, 0.0, 0, c_void_p(after_stream or None))
        if rc == cabi.TOB_E_OOM:
            raise _oom_class()(cabi.last_error())
        if rc != cabi.TOB_OK:
            raise RuntimeError("tob_plan_run_async: " + cabi.last_error())

    def join(self, stream: int) -> None:
        """Makes `stream` (cudaStream_t handle) wait on the device for the run in flight."""
        if cabi.lib.tob_plan_join(self._handle, c_void_p(stream or None)) != cabi.TOB_OK:
            raise RuntimeError("tob_plan_join: " + cabi.last_error())

    def wait(self) -> float:
        out = c_double(0.0)
        if cabi.lib.tob_plan_wait(self._handle, byref(out)) != cabi.TOB_OK:
            raise RuntimeError("tob_plan_wait: " + cabi.last_error())
        return out.value

    def run_interruptible(self, first: int = 0, count: Optional[int] = None, stride: int = 1, target_s: float = 0.25):
        """Same result as `run` (bit for bit: the device accumulator is carried across calls), but
        returns to the interpreter every ~target_s seconds of device work so pending signal handlers —
        the reference's SIGALRM `TimeoutTimer` (src/util/util.py:32-39) — run between chunks of slices."""
        if count is None:
            count = (self.num_slices - first + stride - 1) // stride
        # first chunk: sized from the program's work (flops at ~half the DMMA rate + ~4 us per launch), so a plan whose
        # slices all fit the interval runs in ONE call with both slice lanes busy; later chunks from the measured time
        sf, inv, nl = c_double(0.0), c_double(0.0), ctypes.c_int64(0)
        cabi.lib.tob_plan_work(self._handle, byref(sf), byref(inv), byref(nl))
        est_s = sf.value / 17e12 + nl.value * 4e-6
        acc, done, chunk = 0.0, 0, max(1, min(count, int((target_s - inv.value / 17e12) / est_s))) if est_s > 0 else 1
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


def _locked(method):
    def wrapper(self, *args, **kwargs):
        with self.lock:
            return method(self, *args, **kwargs)

    wrapper.__name__, wrapper.__doc__ = method.__name__, method.__doc__
    return wrapper


class PlanCache:
    """Compiled plans keyed by plan identity (SURVEY.md §8b "Ownership": the backend owns all device memory
    and must release it per call or cache it keyed by plan identity).

    Key: the identity of `plan.tree` and `plan.network` (both held, so the ids stay valid), the slice
    groups' contents and the compile options.  `execution.run`'s OOM retry (src/execution.py:140-142)
    changes `groups_to_slice` and `contract_small` (sliced_execution_plan.py:29-56) replaces `plan.tree`,
    so both miss.  A hit skips the tree walk and the plan compile; the LEAF VALUES are re-read through
    `Tensor.build()` and copied host -> device on every call.  Plans up to 2 GiB (8 GiB in total, least
    recently used first out) stay resident on the device with their captured CUDA graphs; larger ones give
    their arena back after each call and keep only the compiled program."""

    MAX_ENTRIES = 64
    # the same footprint the library's block pool keeps cached anyway (blocks up to 4 GiB, 8 GiB per device): a released
    # arena only moves from the plan to the pool, and the next call pays a full upload (tables, prefix copy) to get it back
    RESIDENT_PLAN_BYTES = 2 << 30
    RESIDENT_TOTAL_BYTES = 8 << 30

    def __init__(self):
        self.entries = OrderedDict()
        self.hits = self.misses = 0
        self.lock = threading.RLock()  # several host threads may contract independent plans through B200API at once

    @staticmethod
    def key(plan, options):
        groups = tuple(tuple(sorted(int(e) for e in g)) for g in plan.groups_to_slice)
        return (id(plan.tree), id(plan.network), groups, options)

    @_locked
    def lookup(self, plan, options):
        k = self.key(plan, options)
        e = self.entries.get(k)
        if e is not None and e["tree"] is plan.tree and e["network"] is plan.network and e["compiled"].busy == 0:
            self.entries.move_to_end(k)
            self.hits += 1
            e["compiled"].busy += 1
            return e
        self.misses += 1
        return None

    @_locked
    def store(self, plan, options, compiled):
        k = self.key(plan, options)
        old = self.entries.get(k)
        if old is not None and old["compiled"].busy > 0:
            return False  # another thread is running this very plan: the caller keeps a private, uncached copy
        if old is not None and old["compiled"] is not compiled:
            self.entries.pop(k)["compiled"].close()
        compiled.busy += 1
        self.entries[k] = {"tree": plan.tree, "network": plan.network, "compiled": compiled}
        for old_key in [q for q, e in self.entries.items() if e["compiled"].busy == 0][: max(0, len(self.entries) - self.MAX_ENTRIES)]:
            self.entries.pop(old_key)["compiled"].close()
        return True

    @_locked
    def drop(self, plan, options):
        e = self.entries.pop(self.key(plan, options), None)
        if e is not None:
            e["compiled"].close()

    @_locked
    def failed(self, compiled):
        compiled.busy = max(0, compiled.busy - 1)
        compiled.release_device()

    @_locked
    def after_call(self, compiled):
        """Residency policy: big arenas go back to the pool at once, small ones stay (LRU within the budget)."""
        compiled.busy = max(0, compiled.busy - 1)
        if not compiled.uploaded:
            return
        if compiled.peak_bytes > self.RESIDENT_PLAN_BYTES:
            compiled.release_device()
            return
        total = 0
        for e in reversed(self.entries.values()):  # most recently used first
            c = e["compiled"]
            if not c.uploaded:
                continue
            total += c.peak_bytes
            if total > self.RESIDENT_TOTAL_BYTES and c is not compiled and c.busy == 0:
                c.release_device()

    @_locked
    def clear(self):
        for e in self.entries.values():
            e["compiled"].close()
        self.entries.clear()


PLAN_CACHE = PlanCache()


class CollectiveOrder:
    """Turnstile for host threads whose `contract_sliced` calls end in a collective.  Under `torch.distributed` every
    rank must issue the count all-reduces of concurrent contractions in the same order; threads do not.  The caller gives
    each call a ticket (the same numbering on every rank: `api.add_argument("collective_ticket", (order, k))`), the
    contractions run concurrently on the device, and the all-reduces pass the turnstile in ticket order 0, 1, 2, ...
    A call that fails before its collective (an invalid plan: the same on every rank) gives its ticket up the same way."""

    def __init__(self, first: int = 0):
        self._cond = threading.Condition()
        self._next = int(first)

    def enter(self, ticket: int) -> None:
        with self._cond:
            if ticket < self._next:
                raise ValueError("collective ticket %d was already served (next is %d)" % (ticket, self._next))
            while self._next != ticket:
                self._cond.wait()

    def leave(self, ticket: int) -> None:
        with self._cond:
            if self._next == ticket:
                self._next = ticket + 1
            self._cond.notify_all()


class B200API:
    """Drop-in for `tensor_network.ALL_APIS[...]` entries (src/tensor_network/__init__.py:12-16).

    Entry types (the table of `NumpyAPI.add_argument`, numpy_apis.py:15-22):
      float64          the DMMA path; counts bit-exact while representable, 1e-9 relative otherwise.
      bigint           exact Python int: residues modulo ~23-bit primes on the same kernels + CRT.
      int / uint       numpy's int64 / uint64 arithmetic wraps modulo 2^64, which is a ring homomorphism of
                       the integers: the exact count is computed as for bigint and reduced modulo 2^64
                       (two's complement for int) — bit-identical to the reference's wrapped result.  Leaf
                       values are truncated toward zero exactly as `numpy.full(shape, w, dtype=int64)` does.
      float32/float16  leaves are rounded to the entry type (as the reference's `create_tensor` does), the
                       contraction runs in float64 on the DMMA path and the count is rounded back to the
                       entry type: at least as accurate as numpy's float32 GEMMs (tolerance vs the reference
                       5e-6 relative for float32, 5e-3 for float16, while the reference stays finite);
                       tensors still occupy 8 bytes per entry on the device, and `get_entry_size` says so."""

    ENTRY_TYPES = ("float64", "float32", "float16", "uint", "int", "bigint")

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
        self._ticket = None  # (CollectiveOrder, k): one-shot, consumed by the next contract_sliced call
        self._mem_limit_bytes = 0
        self._plan_cache = True
        self.last_stats = {}

    # ---- configuration (tensororder.py:205-217, execution.py:82-85) ----
    def add_argument(self, key, value):
        if key == "entry_type":
            if value not in self.ENTRY_TYPES:
                raise ValueError("Unknown b200 type %s" % value)  # numpy_apis.py:26-27
            self._entry_type = value
            # both CLIs make this call right after constructing the backend and before their stopwatch starts
            # (tensororder.py:205-228, execution.py:82-92): the place to create the CUDA context, as the TensorFlow
            # backend initialises its device at import.  No device: nothing happens here, contract_sliced fails loudly.
            cabi.lib.tob_warm(self._resolve_device())
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
        elif key == "collective_ticket":
            if value is not None and not (isinstance(value, tuple) and len(value) == 2 and isinstance(value[0], CollectiveOrder)):
                raise ValueError("collective_ticket takes (CollectiveOrder, int)")
            self._ticket = value
        elif key == "mem_limit_bytes":
            self._mem_limit_bytes = int(value)  # plans needing more raise OutOfMemoryError (=> slice once more)
        elif key == "plan_cache":
            self._plan_cache = bool(value)
        else:
            # same message and exception type as BaseTensorAPI.add_argument (base_api.py:9-12)
            raise ValueError("Invalid argument " + str(key) + " for selected tensor_library")

    def get_entry_size(self):
        return 8  # bytes per entry on the device for every entry type (numpy_apis.py:64-65 for float64)

    def warm(self):
        self._resolve_device()
        if cabi.lib.tob_device_count() <= 0:
            raise RuntimeError("b200 tensor library: no CUDA device")

    # ---- leaf construction: host arrays, like every reference backend (numpy_apis.py:36-40) ----
    _HOST_DTYPES = {"float64": np.float64, "float32": np.float32, "float16": np.float16, "uint": np.uint64,
                    "int": np.int64, "bigint": object}

    def create_tensor(self, shape, default_value=None):
        dtype = self._HOST_DTYPES[self._entry_type]
        if default_value is None:
            return np.empty(shape, dtype=dtype)
        return np.full(shape, default_value, dtype=dtype)

    # ---- plan acquisition: flatten + compile, or a plan-cache hit ----
    def _options_key(self):
        return (self._resolve_device(), self._use_graph, self._kernel_policy, self._hoist, self._microtree,
                self._lanes, self._branches, self._mem_limit_bytes)

    def _cast_leaves(self, leaf_data):
        """What the reference's `create_tensor(shape, value)` does to the leaf values for this entry type."""
        et = self._entry_type
        if et == "float32":
            return leaf_data.astype(np.float32).astype(np.float64)
        if et == "float16":
            return leaf_data.astype(np.float16).astype(np.float64)
        if et in ("int", "uint"):
            if not np.all(np.isfinite(leaf_data)):
                raise ValueError("entry_type %s needs finite weights" % et)
            out = np.trunc(leaf_data)  # numpy float -> int64 conversion truncates toward zero
            if et == "uint" and np.any(out < 0):
                raise ValueError("entry_type uint needs non-negative weights")
            return out
        return leaf_data

    def _acquire(self, execution_plan):
        """Returns (compiled, cache_hit).  The leaf values are always re-read from the caller's tensors."""
        options = self._options_key()
        if self._plan_cache:
            entry = PLAN_CACHE.lookup(execution_plan, options)
            if entry is not None:
                compiled = entry["compiled"]
                data = rebuild_leaf_data(execution_plan, compiled.flat)
                if data is not None:
                    compiled.flat = dataclasses.replace(compiled.flat, leaf_data=self._cast_leaves(np.ascontiguousarray(data)))
                    compiled.cached = True
                    return compiled, True
                compiled.busy = 0
                PLAN_CACHE.drop(execution_plan, options)
        flat = flatten_plan(execution_plan)  # leaves are built through a buffer-backed create_tensor
        flat = dataclasses.replace(flat, leaf_data=self._cast_leaves(flat.leaf_data))
        compiled = CompiledPlan(flat, device=options[0], use_graph=self._use_graph,
                                kernel_policy=self._kernel_policy, hoist_invariant=self._hoist,
                                use_microtree=self._microtree, slice_lanes=self._lanes,
                                dag_branches=self._branches, mem_limit_bytes=self._mem_limit_bytes)
        compiled.cached = bool(self._plan_cache and PLAN_CACHE.store(execution_plan, options, compiled))
        return compiled, False

    def _done_with(self, compiled, failed=False):
        if not getattr(compiled, "cached", True):
            compiled.close()
        elif failed:
            PLAN_CACHE.failed(compiled)
        else:
            PLAN_CACHE.after_call(compiled)

    # ---- the primary entry (base_api.py:17-28, called from execution.py:136) ----
    def contract_sliced(self, execution_plan, num_slice_limit=None):
        ticket, self._ticket = self._ticket, None
        served = [False]
        try:
            return self._contract_sliced(execution_plan, num_slice_limit, ticket, served)
        finally:
            if ticket is not None and not served[0]:  # failed before its collective: give the ticket up
                ticket[0].enter(ticket[1])
                ticket[0].leave(ticket[1])

    def _contract_sliced(self, execution_plan, num_slice_limit, ticket, served):
        if self._entry_type in ("bigint", "int", "uint"):
            if ticket is not None:  # several collectives per call (one per prime): the whole call holds the turnstile
                ticket[0].enter(ticket[1])
                served[0] = True
            try:
                exact = self._contract_exact(execution_plan, num_slice_limit)
            finally:
                if ticket is not None:
                    ticket[0].leave(ticket[1])
            if self._entry_type == "bigint":
                return exact
            wrapped = exact % (1 << 64)  # numpy's integer arithmetic wraps modulo 2^64
            if self._entry_type == "uint":
                return np.uint64(wrapped)
            return np.int64(wrapped - (1 << 64) if wrapped >= (1 << 63) else wrapped)
        t0 = time.perf_counter()
        rank, world = self._rank_world()
        compiled, hit = self._acquire(execution_plan)
        failed = True
        try:
            t1 = time.perf_counter()
            total = compiled.num_slices
            if num_slice_limit is not None:
                total = min(total, int(num_slice_limit))  # itertools.islice(slices, N), base_api.py:23-24
            count = 0 if rank >= total else (total - rank + world - 1) // world
            partial, error = 0.0, None
            try:
                compiled.update_leaves()
                t2 = time.perf_counter()
                partial = compiled.run_interruptible(first=rank if count else 0, count=count, stride=world)
            except BaseException as exc:  # noqa: BLE001  (re-raised below, on every rank)
                error = self._give_up_if_unsliceable(exc, compiled)
                t2 = time.perf_counter()
            t3 = time.perf_counter()
            if world > 1:
                if ticket is not None:
                    ticket[0].enter(ticket[1])
                    served[0] = True
                try:
                    result = self._combine(partial, error)
                finally:
                    if ticket is not None:
                        ticket[0].leave(ticket[1])
            else:
                result = self._raise_or(partial, error)
            failed = False
            self.last_stats = {
                "flatten_compile_s": t1 - t0, "upload_s": t2 - t1, "run_s": t3 - t2,
                "device_ms": compiled.interruptible_stats["device_ms"],
                "launches": compiled.interruptible_stats["launches"],
                "h2d_bytes": int(compiled.flat.leaf_data.nbytes), "d2h_bytes": 32,
                "peak_bytes": compiled.peak_bytes, "slices": total, "rank": rank, "world": world,
                "plan_cache_hit": hit, "t_call": (t0, t1, t2, t3),  # perf_counter: entry, plan acquired, leaves on the device, run done
            }
        finally:
            self._done_with(compiled, failed)
        et = self._entry_type
        with np.errstate(over="ignore"):  # a count beyond the entry type's range is inf there, as in the reference
            return np.float32(result) if et == "float32" else np.float16(result) if et == "float16" else np.float64(result)

    # ---- exact counts (entry_type = bigint, int, uint) ----
    def _contract_exact(self, execution_plan, num_slice_limit=None):
        """Exact integer result, any magnitude: the same compiled plan runs once per prime p < 2^23 with every
        kernel reducing modulo p (all sums stay below 2^53, so FP64 arithmetic is exact), and the residues are
        combined by CRT into the symmetric range (-M/2, M/2].  How many primes: for non-negative tensors one
        float64 pass sizes the answer (no cancellation, so the estimate is good to ~1e-12) and primes are added
        until the reconstruction is stable and agrees with it; with mixed signs the float64 pass can be
        arbitrarily far off, so the count comes from the a-priori bound  |result| <= prod_t ||tensor_t||_1."""
        import math

        t0 = time.perf_counter()
        rank, world = self._rank_world()
        compiled, hit = self._acquire(execution_plan)
        failed = True
        try:
            leaves = compiled.flat.leaf_data
            if not np.all(np.isfinite(leaves)) or not np.array_equal(leaves, np.rint(leaves)) \
                    or np.any(np.abs(leaves) >= 2.0 ** 53):
                raise ValueError("entry_type %s needs integer-valued tensors (weights must be integers)" % self._entry_type)
            total = compiled.num_slices
            if num_slice_limit is not None:
                total = min(total, int(num_slice_limit))
            count = 0 if rank >= total else (total - rank + world - 1) // world
            first = rank if count else 0
            launches = 0

            def one_pass(leaf_data, modulus):
                nonlocal launches
                partial, error = 0.0, None
                try:
                    compiled.update_leaves(leaf_data)
                    compiled.set_modulus(modulus)
                    partial = compiled.run_interruptible(first=first, count=count, stride=world)
                    launches += compiled.interruptible_stats["launches"]
                except BaseException as exc:  # noqa: BLE001
                    error = self._give_up_if_unsliceable(exc, compiled)
                return self._combine(partial, error) if world > 1 else self._raise_or(partial, error)

            flat = compiled.flat
            bound_bits = 2.0  # a-priori: |result| <= product over tensors of their 1-norms
            for l in range(flat.n_leaves):
                off = int(flat.leaf_data_offset[l])
                bound_bits += math.log2(max(float(np.abs(leaves[off: off + (1 << int(flat.leaf_rank[l]))]).sum()), 1.0))
            prime_bits = math.log2(EXACT_PRIMES[-1])
            nonneg = bool(np.all(leaves >= 0))
            estimate = None
            if nonneg:
                estimate = float(one_pass(leaves, 0))
                if not math.isfinite(estimate):
                    estimate = None
            if estimate is not None:
                need = max(1, math.ceil((math.log2(abs(estimate) + 1.0) + 4.0) / prime_bits))
            else:
                need = max(1, math.ceil((bound_bits + 1.0) / prime_bits))  # +1: the symmetric range
            if need > len(EXACT_PRIMES):
                raise OverflowError("count needs more than %d primes" % len(EXACT_PRIMES))
            residues, moduli = [], []
            value, passes = None, 0
            while value is None:
                if len(moduli) >= len(EXACT_PRIMES):
                    raise OverflowError("count needs more than %d primes" % len(EXACT_PRIMES))
                p = EXACT_PRIMES[len(moduli)]
                r = one_pass(np.mod(leaves, p), p)
                passes += 1
                if r != int(r) or not (0 <= r < p * max(world, 1)):
                    raise RuntimeError("exact mode: residue %r is not an integer below the modulus" % r)
                residues.append(int(r) % p)
                moduli.append(p)
                if len(moduli) < need:
                    continue
                m = math.prod(moduli)
                x = crt(residues, moduli)
                if x > m // 2:
                    x -= m  # symmetric range (-M/2, M/2]
                if estimate is None:
                    value = x  # M > 2 * bound: the symmetric representative IS the result
                    break
                # non-negative tensors: stable (one more prime does not change it) and consistent with float64
                if len(moduli) > 1:
                    m1 = math.prod(moduli[:-1])
                    prev = crt(residues[:-1], moduli[:-1])
                    if prev > m1 // 2:
                        prev -= m1
                else:
                    prev = x if abs(x) < m // 4 else None
                if prev == x and abs(x - estimate) <= 1e-6 * abs(estimate) + 1.0:
                    value = x
            failed = False
            self.last_stats = {"exact_passes": passes, "primes": list(moduli), "float_estimate": estimate,
                               "bound_bits": bound_bits, "seconds": time.perf_counter() - t0, "slices": total,
                               "rank": rank, "world": world, "launches": launches, "plan_cache_hit": hit}
        finally:
            try:
                compiled.set_modulus(0)
            except Exception:
                pass
            self._done_with(compiled, failed)
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
            return np.array(value, dtype=self._HOST_DTYPES[self._entry_type])
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
        import sys

        dist = sys.modules.get("torch.distributed")  # never imported => no process group; importing torch costs seconds
        if dist is None:
            return False
        try:
            return dist.is_available() and dist.is_initialized()
        except Exception:
            return False

    def _rank_world(self):
        if self._dist_ready():
            import torch.distributed as dist

            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    @staticmethod
    def _give_up_if_unsliceable(error, compiled):
        """`execution.run` answers OutOfMemoryError by slicing once more, forever (src/execution.py:133-142).  Past
        40 slice groups (10^12 slices) more slicing is not an answer: the leaves, tables and slice-invariant tensors
        alone exceed the limit.  MemoryError ends the loop with "Out of Memory during execution" (execution.py:145-148)."""
        if isinstance(error, (OutOfMemoryError, _oom_class())) and compiled.flat.n_slice_groups >= 40:
            return MemoryError("b200 tensor library: the plan does not fit the device memory limit at any slicing (%s)" % error)
        return error

    @staticmethod
    def _raise_or(value, error):
        if error is not None:
            raise error
        return value

    def _combine(self, partial: float, error) -> float:
        """The one exchange step of the sliced path: an all-reduce (SUM) of [partial count, OOM flag, timeout
        flag, other-error flag].  The flags make failures collective: free memory and SIGALRM timing differ
        per GPU / process, and a rank that left alone would leave its peers blocked in the all-reduce — or,
        after `execution.run` re-sliced only on that rank (src/execution.py:140-142), summing partials of
        different slicings.  Every rank raises the same class, so all of them re-slice (or stop) together."""
        import torch
        import torch.distributed as dist

        oom = error is not None and (isinstance(error, _oom_class()) or isinstance(error, OutOfMemoryError))
        timeout = error is not None and isinstance(error, TimeoutError)
        other = error is not None and not oom and not timeout
        backend = dist.get_backend()
        dev = torch.device("cuda", self._resolve_device()) if backend == "nccl" else torch.device("cpu")
        t = torch.tensor([0.0 if error is not None else partial, float(oom), float(timeout), float(other)],
                         dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total, n_oom, n_timeout, n_other = (float(x) for x in t.tolist())
        if error is not None:
            raise error
        if n_other > 0:
            raise RuntimeError("b200 tensor library: %d other rank(s) failed during contraction" % int(n_other))
        if n_timeout > 0:
            raise TimeoutError("b200 tensor library: %d other rank(s) timed out" % int(n_timeout))
        if n_oom > 0:
            raise _oom_class()("b200 tensor library: %d other rank(s) ran out of device memory" % int(n_oom))
        return total


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
