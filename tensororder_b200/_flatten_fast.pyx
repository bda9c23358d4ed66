# cython: language_level=3, boundscheck=False, wraparound=False, cdivision=True
# distutils: language = c++
"""Compiled hot loop of `flatten_plan` (tensororder_b200/flatten.py has the same logic in Python and the
reference citations).  The reference builds its own Cython (`src/setup.py`); this is the same kind of thin
layer between its Python objects and the C ABI: one pass over `plan.tree.iterate_postorder()` appending to
C++ vectors, and a leaf factory that hands `Tensor.build()` views of ONE growing float64 buffer (created with
the numpy C API, no slicing/reshape calls), which is the buffer `tob_plan_upload` receives."""
import numpy as np

cimport numpy as cnp
from cpython.ref cimport Py_INCREF
from libc.stdint cimport int32_t, int64_t
from libc.string cimport memcpy
from libcpp.vector cimport vector

cnp.import_array()


cdef class LeafArena:
    cdef public object buf      # numpy float64 array owning the storage
    cdef double* ptr
    cdef Py_ssize_t cap
    cdef public Py_ssize_t used
    cdef object last

    def __init__(self, Py_ssize_t capacity=4096):
        self.buf = np.empty(capacity, dtype=np.float64)
        self.ptr = <double*> cnp.PyArray_DATA(<cnp.ndarray> self.buf)
        self.cap = capacity
        self.used = 0
        self.last = None

    cdef void grow(self, Py_ssize_t need):
        cdef Py_ssize_t cap = max(2 * self.cap, need)
        grown = np.empty(cap, dtype=np.float64)
        cdef double* gp = <double*> cnp.PyArray_DATA(<cnp.ndarray> grown)
        memcpy(gp, self.ptr, self.used * sizeof(double))
        self.buf = grown  # views handed out earlier keep the old array alive; their data is copied already
        self.ptr = gp
        self.cap = cap

    def factory(self, shape, default_value=None):
        cdef cnp.npy_intp dims[32]
        cdef int nd = len(shape)
        cdef Py_ssize_t n = 1
        cdef int i
        if nd > 32:
            raise ValueError("tensor rank above 32")
        for i in range(nd):
            dims[i] = shape[i]
            n *= dims[i]
        if self.used + n > self.cap:
            self.grow(self.used + n)
        cdef cnp.ndarray view = cnp.PyArray_SimpleNewFromData(nd, dims, cnp.NPY_FLOAT64, <void*> (self.ptr + self.used))
        Py_INCREF(self.buf)
        cnp.PyArray_SetBaseObject(view, self.buf)
        cdef double v
        cdef double* p
        if default_value is not None:
            v = default_value
            p = self.ptr + self.used
            for i in range(n):
                p[i] = v
        self.last = view
        return view

    cdef Py_ssize_t commit(self, object built, Py_ssize_t n) except -1:
        """Keep `built` (normally the view just handed out) as the next n doubles of the buffer."""
        cdef Py_ssize_t offset = self.used
        cdef cnp.ndarray src
        if built is not self.last:  # build() returned something of its own: copy it in
            src = np.ascontiguousarray(built, dtype=np.float64).reshape(-1)
            if src.shape[0] != n:
                raise ValueError("leaf size mismatch")
            if offset + n > self.cap:
                self.grow(offset + n)
            memcpy(self.ptr + offset, cnp.PyArray_DATA(src), n * sizeof(double))
        self.used = offset + n
        return offset

    def data(self):
        return self.buf[: self.used]


cdef object _as_array(vector[int32_t]& v):
    cdef cnp.npy_intp n = v.size()
    cdef cnp.ndarray out = np.empty(n, dtype=np.int32)
    if n:
        memcpy(cnp.PyArray_DATA(out), v.data(), n * sizeof(int32_t))
    return out


cdef object _as_array64(vector[int64_t]& v):
    cdef cnp.npy_intp n = v.size()
    cdef cnp.ndarray out = np.empty(n, dtype=np.int64)
    if n:
        memcpy(cnp.PyArray_DATA(out), v.data(), n * sizeof(int64_t))
    return out


def flatten_tree(plan, dict group_of):
    """Returns (node_left, node_right, node_leaf, leaf_rank, leaf_data_offset, leaf_axis_start, leaf_axis_edge,
    leaf_data, leaf_tensor_index).  `group_of`: edge id -> slice group index (non-empty groups only)."""
    cdef vector[int32_t] node_left, node_right, node_leaf, leaf_rank, axis_start, axis_edge, stack
    cdef vector[int64_t] leaf_off
    cdef LeafArena arena = LeafArena()
    cdef dict built_at = {}
    cdef list leaf_tensor_index = []
    cdef bint sliced = len(group_of) > 0
    cdef Py_ssize_t size, off
    cdef int rank, pos = 0, n_leaves = 0, right, e_i
    cdef object network = plan.network
    cdef object index_list = network.index_list
    cdef object factory = arena.factory
    cdef object t, edges, e, built, g, cached
    axis_start.push_back(0)
    for node in plan.tree.iterate_postorder():
        if node.is_leaf:
            t = node.tensor_index
            edges = index_list(t)
            rank = len(edges)
            cached = built_at.get(t)
            if cached is None:
                size = (<Py_ssize_t> 1) << rank
                built = network[t].build(factory)
                if getattr(built, "size", size) != size:
                    raise ValueError("tensor %d: only indices of extent 2 are supported" % t)
                off = arena.commit(built, size)
                built_at[t] = off
            else:
                off = cached
            node_left.push_back(-1)
            node_right.push_back(-1)
            node_leaf.push_back(n_leaves)
            n_leaves += 1
            leaf_rank.push_back(rank)
            leaf_off.push_back(off)
            for e in edges:
                e_i = e
                if e_i < 0:
                    raise ValueError("tensor %d has a dangling index; the network cannot contract to a scalar" % t)
                if sliced:
                    g = group_of.get(e)
                    axis_edge.push_back(e_i if g is None else -(<int> g + 1))
                else:
                    axis_edge.push_back(e_i)
            axis_start.push_back(<int32_t> axis_edge.size())
            leaf_tensor_index.append(t)
        else:
            if stack.size() < 2:
                raise ValueError("contraction tree is not a single rooted tree")
            right = stack.back()
            stack.pop_back()
            node_left.push_back(stack.back())
            stack.pop_back()
            node_right.push_back(right)
            node_leaf.push_back(-1)
        stack.push_back(pos)
        pos += 1
    if stack.size() != 1:
        raise ValueError("contraction tree is not a single rooted tree")
    return (_as_array(node_left), _as_array(node_right), _as_array(node_leaf), _as_array(leaf_rank),
            _as_array64(leaf_off), _as_array(axis_start), _as_array(axis_edge), arena.data(), leaf_tensor_index)


def rebuild_leaves(network, list leaf_tensor_index, cnp.ndarray leaf_rank, cnp.ndarray leaf_off, Py_ssize_t total):
    """Plan-cache hits: the leaf values re-read through `Tensor.build()` in the order `flatten_tree` stored
    them; None when they no longer fit the cached structure (see flatten.rebuild_leaf_data)."""
    cdef LeafArena arena = LeafArena(max(total, 16))
    cdef object factory = arena.factory
    cdef set seen = set()
    cdef Py_ssize_t j, size
    cdef object t, built
    cdef int32_t* rk = <int32_t*> cnp.PyArray_DATA(leaf_rank)
    cdef int64_t* lo = <int64_t*> cnp.PyArray_DATA(leaf_off)
    for j in range(len(leaf_tensor_index)):
        t = leaf_tensor_index[j]
        if t in seen:
            continue
        seen.add(t)
        size = (<Py_ssize_t> 1) << rk[j]
        if arena.used != lo[j]:
            return None
        built = network[t].build(factory)
        if getattr(built, "size", size) != size:
            return None
        arena.commit(built, size)
    if arena.used != total:
        return None
    return arena.data()
