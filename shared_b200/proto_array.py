"""Host-side mirror of the reference's structural array operations -- the CALLERS of ArrayKernel.map / slice.

`ProtoArray` follows org.shared.array.ProtoArray (src/org/shared/array/ProtoArray.java:56-527) and the slice helpers of
org.shared.array.ArrayBase (ArrayBase.java:81-215): an array is (flat values, indexing order, dims, strides), and tile /
transpose / shift / subarray / reverse / concat / reverseOrder / the slice variants are each ONE `map` or `slice` call
with bounds computed here exactly as the reference computes them. The class is generic over the kernel provider: it is
handed any object with the SPI method names (`CudaArrayKernel` in production; the parity tests hand it the reference's
native kernel to pin the bounds arithmetic to the literal known-answer arrays of RealArrayTest / IntegerArrayTest on a
box without a GPU). Values are float64 (real, or interleaved complex with a trailing dimension of 2 like
ComplexArray) or int32. No array contents are computed in this module -- everything goes through the kernel.
"""
import numpy as np

FAR, NEAR = "FAR", "NEAR"


def order_strides(order, dims):
    """IndexingOrder.strides (Array.java:240-287): FAR = row-major, NEAR = column-major."""
    n = len(dims)
    s = [1] * n
    if order == FAR:
        for i in range(n - 2, -1, -1):
            s[i] = s[i + 1] * dims[i + 1]
    else:
        for i in range(1, n):
            s[i] = s[i - 1] * dims[i - 1]
    return s


def canonicalize_slices(src_slices, dst_slices):
    """ArrayBase.canonicalizeSlices(int[][], int[], int[][], int[]) (ArrayBase.java:81-113)."""
    if len(src_slices) != len(dst_slices):
        raise ValueError("Dimensionality mismatch")
    out = []
    for dim, (s, d) in enumerate(zip(src_slices, dst_slices)):
        if len(s) != len(d):
            raise ValueError("Dimension mismatch")
        for a, b in zip(s, d):
            out += [int(a), int(b), dim]
    return out


def reverse_slices(dims, op_dims):
    """ArrayBase.createReverseSlices (ArrayBase.java:176-215)."""
    out = []
    for dim, size in enumerate(dims):
        for j in range(size):
            out += [j, size - 1 - j if dim in op_dims else j, dim]
    return out


class ProtoArray:
    def __init__(self, kernel, values=None, dims=(), order=FAR, strides=None, dtype=np.float64):
        self.kernel = kernel
        self.order = order or FAR  # a literal without an order is FAR (RealArray.java's default)
        self.dims = [int(d) for d in dims]
        self.strides = list(strides) if strides is not None else order_strides(self.order, self.dims)
        n = int(np.prod(self.dims)) if self.dims else 0
        if values is None:
            self.values = np.zeros(n, dtype=dtype)
        else:
            self.values = np.ascontiguousarray(values, dtype=np.asarray(values).dtype if isinstance(values, np.ndarray) else dtype)
            if self.values.size != n:
                raise ValueError("Invalid dimensions and/or strides")

    def _wrap(self, dims, order=None, strides=None):
        order = order or self.order
        return ProtoArray(self.kernel, None, dims, order, strides, self.values.dtype)

    def size(self, d):
        return self.dims[d]

    # ---- the two kernel calls (ProtoArray.java:144-170) ------------------------------------------------------------
    def map(self, dst, *bounds):
        if dst is self:
            raise ValueError("Source and destination cannot be the same")
        self.kernel.map(np.asarray(bounds, dtype=np.int32), self.values, self.dims, self.strides, dst.values, dst.dims,
                        dst.strides)
        return dst

    def splice(self, dst, *slices):
        if dst is self:
            raise ValueError("Source and destination cannot be the same")
        self.kernel.slice(np.asarray(slices, dtype=np.int32), self.values, self.dims, self.strides, dst.values, dst.dims,
                          dst.strides)
        return dst

    # ---- slice variants (ProtoArray.java:174-243) -----------------------------------------------------------------------
    def slice_into(self, src_slices, dst, dst_slices):
        """slice(int[][] srcSlices, T dst, int[][] dstSlices)"""
        return self.splice(dst, *canonicalize_slices(src_slices, dst_slices))

    def slice_to(self, dst, *dst_slices):
        """slice(T dst, int[]... dstSlices): every source index, in order, to the listed destination indices."""
        if len(dst_slices) != len(self.dims) or any(len(s) != d for s, d in zip(dst_slices, self.dims)):
            raise ValueError("Dimension mismatch")
        return self.splice(dst, *canonicalize_slices([range(d) for d in self.dims], dst_slices))

    def slice(self, *src_slices):
        """slice(int[]... srcSlices): a new array holding the listed source indices."""
        if len(src_slices) != len(self.dims):
            raise ValueError("Dimensionality mismatch")
        dst = self._wrap([len(s) for s in src_slices])
        return self.splice(dst, *canonicalize_slices(src_slices, [range(len(s)) for s in src_slices]))

    def slice_fill(self, value, *src_slices):
        """slice(E value, int[]... srcSlices): assigns `value` to the cross product of the listed indices of THIS array,
        in place -- the reference builds a value-filled array of that shape and slices it into `this`
        (ProtoArray.java:207-219)."""
        w = self._wrap([len(s) for s in src_slices])
        w.values[:] = value
        return w.slice_to(self, *src_slices)

    # ---- structural operations: one map / slice call each (ProtoArray.java:246-484) -----------------------------------
    def tile(self, *repetitions):
        if len(repetitions) != len(self.dims):
            raise ValueError("Dimensionality mismatch")
        dst = self._wrap([d * r for d, r in zip(self.dims, repetitions)])
        bounds = [v for d in dst.dims for v in (0, 0, d)]  # size > dimension: the source wraps around
        return self.map(dst, *bounds)

    def transpose(self, *permutation):
        n = len(self.dims)
        if len(permutation) != n:
            raise ValueError("Dimensionality mismatch")
        if sorted(permutation) != list(range(n)):
            raise ValueError("Invalid permutation")
        new_dims = [0] * n
        for dim in range(n):
            new_dims[permutation[dim]] = self.dims[dim]
        dst = self._wrap(new_dims)
        bounds = [v for d in self.dims for v in (0, 0, d)]
        # the destination is addressed with the SOURCE's dims and its own strides permuted
        self.kernel.map(np.asarray(bounds, dtype=np.int32), self.values, self.dims, self.strides, dst.values, self.dims,
                        [dst.strides[permutation[dim]] for dim in range(n)])
        return dst

    def shift(self, *shifts):
        if len(shifts) != len(self.dims):
            raise ValueError("Dimensionality mismatch")
        dst = self._wrap(self.dims, strides=self.strides)
        bounds = [v for d, s in zip(self.dims, shifts) for v in (0, s, d)]
        return self.map(dst, *bounds)

    def subarray(self, *bounds):
        n = len(self.dims)
        if len(bounds) != 2 * n:
            raise ValueError("Invalid subarray bounds")
        new_dims = [bounds[2 * d + 1] - bounds[2 * d] for d in range(n)]
        dst = self._wrap(new_dims)
        return self.map(dst, *[v for d in range(n) for v in (bounds[2 * d], 0, new_dims[d])])

    def reverse(self, *op_dims):
        dst = self._wrap(self.dims, strides=self.strides)
        return self.splice(dst, *reverse_slices(self.dims, op_dims))

    def reshape(self, *dims):
        if int(np.prod(dims)) != int(np.prod(self.dims)):
            raise ValueError("Cardinality mismatch")
        dst = self._wrap(list(dims))
        dst.values[:] = self.values  # System.arraycopy of the backing array (ProtoArray.java:395)
        return dst

    def concat(self, op_dim, *srcs):
        n = len(self.dims)
        if not 0 <= op_dim < n:
            raise ValueError("Invalid dimension")
        parts = (self,) + tuple(srcs)
        for p in parts:
            if any(self.dims[d] != p.dims[d] for d in range(n) if d != op_dim):
                raise ValueError("Dimension mismatch")
        new_dims = list(self.dims)
        new_dims[op_dim] = sum(p.dims[op_dim] for p in parts)
        dst = self._wrap(new_dims)
        bounds = [v for d in self.dims for v in (0, 0, d)]
        offset = 0
        for p in parts:
            bounds[3 * op_dim + 1] = offset
            bounds[3 * op_dim + 2] = p.dims[op_dim]
            p.map(dst, *bounds)
            offset += p.dims[op_dim]
        return dst

    def reverseOrder(self):
        new_order = NEAR if self.order == FAR else FAR
        dst = self._wrap(self.dims, order=new_order)
        return self.map(dst, *[v for d in self.dims for v in (0, 0, d)])

    def clone(self):
        dst = self._wrap(self.dims, strides=self.strides)
        dst.values[:] = self.values
        return dst

    # ---- Matrix.mMul for real matrices (RealArray.java:104-125): row-major operands only ----------------------------------
    def mMul(self, other):
        if self.order != FAR or other.order != FAR or len(self.dims) != 2 or len(other.dims) != 2:
            raise ValueError("Matrices must be two-dimensional and row-major")
        if self.dims[1] != other.dims[0]:
            raise ValueError("Dimensionality mismatch")
        dst = self._wrap([self.dims[0], other.dims[1]])
        self.kernel.mul(self.values, other.values, self.dims[0], other.dims[1], dst.values, False)
        return dst
