// sparse_core.cuh -- ArrayKernel.insertSparse / sliceSparse, replacing
// native/src/shared/SparseOps.cpp:26-290, SparseOpsInsert.cpp:26-168 and SparseOpsSlice.cpp:26-589 (Java twin:
// src/org/shared/array/kernel/SparseOps.java).
//
// A sparse array is (values, sorted unique physical indices, and per dimension a stable ordering of the elements by
// their logical index in that dimension: `indirections`, with the bucket starts in `indirectionOffsets`). Both
// operations end in the same step -- merge a sorted "old" list with a sorted "new" list (new wins on equal indices)
// and rebuild the per-dimension metadata -- which the reference does with sequential two-pointer merges, counting
// sorts and std::sort. Here every step is data-parallel:
//   merge      rank of an element in the union = own position + lower_bound in the other list - number of matched
//              pairs below it (one exclusive scan over match flags);
//   metadata   per dimension a stable sort of (logical index, position) pairs -- all dimensions in one segmented sort
//              -- and one lower_bound per (dimension, logical index) for the bucket starts;
//   insert     sort of the (index, position) pairs of the new elements, duplicate check on neighbours;
//   slice      per source element the number of destination cells it maps to (product of per-dimension lookup
//              counts), exclusive scan, one thread per produced (destination index, source element) pair, sort,
//              first-of-run selection (the reference's stable sort + dedupe keeps the lowest source position), and
//              a flag/scan/compact of the destination elements the slice does not overwrite.
// Results are bit-exact (integer / index / data movement).
//
// The pipeline is written against an executor (for_each / scan / sort_pairs / scratch / synchronising reads):
// array_sparse.cu supplies the CUDA one (grid-stride kernels, a three-phase scan, the bitonic pair sort of
// array_dimension.cu); tests/emu/sparse_emu.cpp supplies a serial host one, so that the same orchestration and the
// same per-element code are checked against the reference's native SparseOps on the CPU.
#pragma once

#include <limits.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <stdexcept>
#include <vector>

#ifdef __CUDACC__
#define SSTC_HD __host__ __device__
#else
#define SSTC_HD
#endif

namespace sstc {

constexpr int kSparseMaxDims = 16;

enum : int { SP_ERR_PHYSICAL = 0, SP_ERR_DUPLICATE = 1, SP_ERR_OVERFLOW = 2, SP_ERR_SLOTS = 4 };

// scratch slots of the executor
enum : int {
    SPB_ERR = 0, SPB_OLD_I, SPB_NEW_I, SPB_OLD_V, SPB_NEW_V, SPB_OLD_SRC, SPB_NEW_SRC, SPB_KEYS, SPB_PAY, SPB_FLAG,
    SPB_SCAN, SPB_OLD_ASSIGN, SPB_NEW_ASSIGN, SPB_OUT_V, SPB_OUT_I, SPB_OUT_IO, SPB_OUT_II, SPB_MAPLEN, SPB_OFF,
    SPB_SRC_OF, SPB_RES_I, SPB_RES_SRC, SPB_KEPT_I, SPB_KEPT_SRC, SPB_TABLE_A, SPB_TABLE_B, SPB_TABLE_C, SPB_TABLE_D,
    SPB_SORTED_I, SPB_PERM, SPB_SLOTS
};

struct SparseShape {
    int nd;
    int32_t dims[kSparseMaxDims], strides[kSparseMaxDims], dim_offsets[kSparseMaxDims + 1];
    int32_t prod;       // product of dims
    int32_t n_offsets;  // sum of dims + nd
};

// What the C ABI hands back: host copies of the four arrays of org.shared.array.sparse.SparseArrayState.
struct SparseHostResult {
    int32_t count = 0;
    std::vector<unsigned char> values;
    std::vector<int32_t> indices, offsets, indirections;
};

SSTC_HD inline uint64_t sp_key(int32_t v) { return (uint64_t) ((int64_t) v + 2147483648LL); }
SSTC_HD inline int32_t sp_unkey(uint64_t k) { return (int32_t) ((int64_t) k - 2147483648LL); }

SSTC_HD inline int32_t sp_lower_bound(const int32_t *a, int32_t n, int32_t v) {  // first position with a[pos] >= v
    int32_t lo = 0, hi = n;
    while (lo < hi) {
        const int32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) {
            lo = mid + 1;
        } else {
            hi = mid;
        }
    }
    return lo;
}

SSTC_HD inline int32_t sp_lower_bound_key(const uint64_t *a, int32_t n, uint64_t v) {
    int32_t lo = 0, hi = n;
    while (lo < hi) {
        const int32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) {
            lo = mid + 1;
        } else {
            hi = mid;
        }
    }
    return lo;
}

SSTC_HD inline int64_t sp_upper_bound(const int64_t *a, int64_t n, int64_t v) {  // first position with a[pos] > v
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] <= v) {
            lo = mid + 1;
        } else {
            hi = mid;
        }
    }
    return lo;
}

// SparseOps.cpp:204-214: logical index per dimension by successive division; false when the physical index or one of
// the quotients is out of range (the reference then walks off its tables).
SSTC_HD inline bool sp_logical(const SparseShape &s, int32_t phys, int32_t *logical) {
    if (!(phys >= 0 && phys < s.prod)) {
        return false;
    }
    int32_t acc = phys;
    bool ok = true;
    for (int d = 0; d < s.nd; d++) {
        const int32_t q = acc / s.strides[d];
        acc %= s.strides[d];
        logical[d] = q;
        ok = ok && q < s.dims[d];
    }
    return ok;
}

// ---- per-element functors (one call per index; any order, any parallelism) --------------------------------------------

struct SpIota {
    int32_t *out;
    SSTC_HD void operator()(int64_t i) const { out[i] = (int32_t) i; }
};

// insert: (biased index, position) pairs of the new elements, padded to a power of two with sentinels
struct SpNewKeys {
    const int32_t *new_i;
    int64_t n;
    uint64_t *keys;
    int32_t *pay;
    SSTC_HD void operator()(int64_t i) const {
        keys[i] = i < n ? sp_key(new_i[i]) : ~0ULL;
        pay[i] = i < n ? (int32_t) i : INT_MAX;
    }
};

struct SpUnpackNew {
    const uint64_t *keys;
    const int32_t *pay;
    int32_t *sorted_i, *perm, *err;
    SSTC_HD void operator()(int64_t r) const {
        sorted_i[r] = sp_unkey(keys[r]);
        perm[r] = pay[r];
        if (r > 0 && keys[r] == keys[r - 1]) {
            err[SP_ERR_DUPLICATE] = 1;  // SparseOpsInsert.cpp:130-135
        }
    }
};

struct SpDupFlag {
    const int32_t *old_i;
    int32_t n_old;
    const int32_t *new_i;
    int32_t *flag;
    SSTC_HD void operator()(int64_t k) const {
        const int32_t v = new_i[k];
        const int32_t p = sp_lower_bound(old_i, n_old, v);
        flag[k] = (p < n_old && old_i[p] == v) ? 1 : 0;
    }
};

// SparseOps.cpp:145-178 without the sequential walk: slot of new element k / old element j in the merged list
struct SpAssignNew {
    const int32_t *old_i;
    int32_t n_old;
    const int32_t *new_i;
    const int64_t *dup_before;
    int32_t *assign;
    SSTC_HD void operator()(int64_t k) const {
        assign[k] = (int32_t) (k + sp_lower_bound(old_i, n_old, new_i[k]) - dup_before[k]);
    }
};

struct SpAssignOld {
    const int32_t *new_i;
    int32_t n_new;
    const int32_t *old_i;
    const int64_t *dup_before;
    int32_t *assign;
    SSTC_HD void operator()(int64_t j) const {
        const int32_t lb = sp_lower_bound(new_i, n_new, old_i[j]);
        assign[j] = (int32_t) (j + lb - dup_before[lb]);
    }
};

template <class U>
struct SpScatter {  // SparseOps.cpp:48-54 (values) and :189-195 (indices)
    const int32_t *assign, *idx, *src_pos;
    const U *vals;
    int32_t *out_i;
    U *out_v;
    SSTC_HD void operator()(int64_t i) const {
        const int32_t a = assign[i];
        out_i[a] = idx[i];
        out_v[a] = vals[src_pos[i]];
    }
};

// per dimension: (logical index, position) for every merged element, padded; also the range check of SparseOps.cpp:206
struct SpDimKeys {
    SparseShape shape;
    const int32_t *indices;
    int64_t count, npad;
    uint64_t *keys;
    int32_t *pay, *err;
    SSTC_HD void operator()(int64_t i) const {
        const int64_t d = i / npad, r = i - d * npad;
        if (r < count) {
            int32_t logical[kSparseMaxDims];
            int32_t key = 0;
            if (sp_logical(shape, indices[r], logical)) {
                key = logical[d];
            } else {
                err[SP_ERR_PHYSICAL] = 1;
            }
            keys[i] = (uint64_t) (uint32_t) key;
            pay[i] = (int32_t) r;
        } else {
            keys[i] = ~0ULL;
            pay[i] = INT_MAX;
        }
    }
};

struct SpIndirections {  // SparseOps.cpp:236-248
    const int32_t *pay;
    int64_t count, npad;
    int32_t *out;
    SSTC_HD void operator()(int64_t i) const {
        const int64_t d = i / count, r = i - d * count;
        out[i] = pay[d * npad + r];
    }
};

struct SpOffsets {  // SparseOps.cpp:218-232
    SparseShape shape;
    const uint64_t *keys;
    int64_t count, npad;
    int32_t *out;
    SSTC_HD void operator()(int64_t i) const {
        int d = 0;
        while (d + 1 < shape.nd && i >= shape.dim_offsets[d + 1]) {
            d++;
        }
        const int32_t v = (int32_t) (i - shape.dim_offsets[d]);
        out[i] = v >= shape.dims[d] ? (int32_t) count
                                    : sp_lower_bound_key(keys + d * npad, (int32_t) count, (uint64_t) (uint32_t) v);
    }
};

// slice: number of destination cells source element e lands in (SparseOpsSlice.cpp:262-291)
struct SpMapLen {
    SparseShape src;
    const int32_t *src_i, *lookup_counts;
    int32_t *maplen, *err;
    SSTC_HD void operator()(int64_t e) const {
        int32_t logical[kSparseMaxDims];
        if (!sp_logical(src, src_i[e], logical)) {
            err[SP_ERR_PHYSICAL] = 1;
            maplen[e] = 0;
            return;
        }
        int64_t len = 1;
        for (int d = 0; d < src.nd; d++) {
            len *= lookup_counts[src.dim_offsets[d] + logical[d]];
            if (len > INT_MAX) {
                err[SP_ERR_OVERFLOW] = 1;
                len = 0;
            }
        }
        maplen[e] = (int32_t) len;
    }
};

// slice: produced pair t -> (destination physical index, source element), SparseOpsSlice.cpp:326-368
struct SpExpand {
    SparseShape src;
    int32_t dst_strides[kSparseMaxDims];
    const int32_t *src_i, *lookup_offsets, *lookup_counts, *dst_lookups;
    const int64_t *off;
    int64_t n_src, total;
    uint64_t *keys;
    int32_t *pay, *src_of;
    SSTC_HD void operator()(int64_t t) const {
        if (t >= total) {
            keys[t] = ~0ULL;
            pay[t] = INT_MAX;
            return;
        }
        const int64_t e = sp_upper_bound(off, n_src + 1, t) - 1;
        int64_t r = t - off[e];
        int32_t logical[kSparseMaxDims];
        sp_logical(src, src_i[e], logical);
        int32_t idx = 0;
        for (int d = src.nd - 1; d >= 0; d--) {
            const int32_t slot = src.dim_offsets[d] + logical[d];
            const int32_t size = lookup_counts[slot];
            const int32_t choice = (int32_t) (r % size);
            r /= size;
            idx += dst_strides[d] * dst_lookups[lookup_offsets[slot] + choice];
        }
        keys[t] = sp_key(idx);
        pay[t] = (int32_t) t;
        src_of[t] = (int32_t) e;
    }
};

struct SpFirstFlag {  // SparseOpsSlice.cpp:565-580: of equal destination indices the first of the stable order stays
    const uint64_t *keys;
    int32_t *flag;
    SSTC_HD void operator()(int64_t r) const { flag[r] = (r == 0 || keys[r] != keys[r - 1]) ? 1 : 0; }
};

struct SpCompactNew {
    const uint64_t *keys;
    const int32_t *pay, *src_of, *flag;
    const int64_t *pos;
    int32_t *res_i, *res_src;
    SSTC_HD void operator()(int64_t r) const {
        if (flag[r]) {
            res_i[pos[r]] = sp_unkey(keys[r]);
            res_src[pos[r]] = src_of[pay[r]];
        }
    }
};

// slice: destination elements that survive = not inside the destination slice region in EVERY dimension
// (SparseOpsSlice.cpp:236-241 + 372-404)
struct SpKeepFlag {
    SparseShape dst;
    const int32_t *dst_i, *member;
    int32_t *flag, *err;
    SSTC_HD void operator()(int64_t j) const {
        int32_t logical[kSparseMaxDims];
        if (!sp_logical(dst, dst_i[j], logical)) {
            err[SP_ERR_PHYSICAL] = 1;
            flag[j] = 1;
            return;
        }
        bool inside = true;
        for (int d = 0; d < dst.nd; d++) {
            inside = inside && member[dst.dim_offsets[d] + logical[d]] != 0;
        }
        flag[j] = inside ? 0 : 1;
    }
};

struct SpCompactOld {
    const int32_t *dst_i, *flag;
    const int64_t *pos;
    int32_t *kept_i, *kept_src;
    SSTC_HD void operator()(int64_t j) const {
        if (flag[j]) {
            kept_i[pos[j]] = dst_i[j];
            kept_src[pos[j]] = (int32_t) j;
        }
    }
};

// ---- host-side pieces: validation and the small per-call tables ----------------------------------------------------------

inline int64_t sp_next_pow2(int64_t n) {
    int64_t p = 1;
    while (p < n) {
        p <<= 1;
    }
    return p;
}

// MappingOps::checkDimensions (MappingOps.cpp:73-91) + the dimension-offset check of SparseOpsInsert.cpp:103-108
inline SparseShape sp_make_shape(const int32_t *D, const int32_t *S, const int32_t *Do, int nd, bool has_elements) {
    if (nd < 0 || nd > kSparseMaxDims) {
        throw std::runtime_error("Invalid arguments");
    }
    SparseShape s;
    memset(&s, 0, sizeof(s));
    s.nd = nd;
    int64_t prod = 1, acc = 0, sum = 0;
    for (int d = 0; d < nd; d++) {
        if (D[d] < 0 || S[d] < 0) {
            throw std::runtime_error("Invalid dimensions and/or strides");
        }
        prod *= D[d];
        if (prod > INT_MAX) {
            throw std::runtime_error("Invalid dimensions and/or strides");
        }
        acc += (int64_t) (D[d] - 1) * S[d];
        sum += D[d];
        s.dims[d] = D[d];
        s.strides[d] = S[d];
    }
    if (acc != prod - 1) {
        throw std::runtime_error("Invalid dimensions and/or strides");
    }
    for (int d = 0; d < nd; d++) {
        if (Do[d + 1] - Do[d] - 1 != D[d]) {
            throw std::runtime_error("Invalid arguments");
        }
    }
    // the reference indexes its tables with Do[] as given and divides by the strides as given: anything but offsets
    // starting at 0 and non-zero strides makes it read or write out of bounds / trap
    if (nd > 0 && Do[0] != 0) {
        throw std::runtime_error("Invalid arguments");
    }
    for (int d = 0; d < nd; d++) {
        if (S[d] == 0 && has_elements) {
            throw std::runtime_error("Invalid dimensions and/or strides");
        }
        s.dim_offsets[d] = Do[d];
    }
    s.dim_offsets[nd] = nd > 0 ? Do[nd] : 0;
    s.prod = (int32_t) prod;
    s.n_offsets = (int32_t) (sum + nd);
    return s;
}

template <class Exec>
inline void sp_check_errors(Exec &ex, const int32_t *err) {
    int32_t h[SP_ERR_SLOTS];
    ex.download(h, err, sizeof(h));
    ex.sync();
    if (h[SP_ERR_DUPLICATE]) {
        throw std::runtime_error("Duplicate values are not allowed");
    }
    if (h[SP_ERR_PHYSICAL]) {
        throw std::runtime_error("Invalid physical index");
    }
    if (h[SP_ERR_OVERFLOW]) {
        throw std::runtime_error("Invalid arguments");
    }
}

// ---- the shared tail: merge two sorted index lists and rebuild the metadata (SparseOps.cpp:129-260, 26-127) -------------

template <class U, class Exec>
inline void sp_merge(Exec &ex, const SparseShape &shape, int32_t *err,
        const int32_t *old_i, const int32_t *old_src, int32_t n_old, const U *old_v,
        const int32_t *new_i, const int32_t *new_src, int32_t n_new, const U *new_v, SparseHostResult &res) {
    int32_t *flag = ex.template buf<int32_t>(SPB_FLAG, (int64_t) n_new + 1);
    int64_t *dup_before = ex.template buf<int64_t>(SPB_SCAN, (int64_t) n_new + 1);
    ex.for_each(n_new, SpDupFlag{old_i, n_old, new_i, flag});
    ex.scan(flag, n_new, dup_before);
    const int64_t dups = ex.read(dup_before + n_new);
    const int64_t count = (int64_t) n_old + n_new - dups;
    if (count > INT_MAX) {
        throw std::runtime_error("Invalid arguments");
    }
    int32_t *old_assign = ex.template buf<int32_t>(SPB_OLD_ASSIGN, n_old);
    int32_t *new_assign = ex.template buf<int32_t>(SPB_NEW_ASSIGN, n_new);
    ex.for_each(n_old, SpAssignOld{new_i, n_new, old_i, dup_before, old_assign});
    ex.for_each(n_new, SpAssignNew{old_i, n_old, new_i, dup_before, new_assign});
    U *out_v = ex.template buf<U>(SPB_OUT_V, count);
    int32_t *out_i = ex.template buf<int32_t>(SPB_OUT_I, count);
    ex.for_each(n_old, SpScatter<U>{old_assign, old_i, old_src, old_v, out_i, out_v});
    ex.for_each(n_new, SpScatter<U>{new_assign, new_i, new_src, new_v, out_i, out_v});  // after the old ones: new wins
    // metadata: every dimension's stable ordering by logical index in one segmented sort
    const int nd = shape.nd;
    const int64_t npad = sp_next_pow2(count);
    uint64_t *keys = ex.template buf<uint64_t>(SPB_KEYS, (int64_t) nd * npad);
    int32_t *pay = ex.template buf<int32_t>(SPB_PAY, (int64_t) nd * npad);
    int32_t *out_io = ex.template buf<int32_t>(SPB_OUT_IO, shape.n_offsets);
    int32_t *out_ii = ex.template buf<int32_t>(SPB_OUT_II, (int64_t) nd * count);
    ex.for_each((int64_t) nd * npad, SpDimKeys{shape, out_i, count, npad, keys, pay, err});
    ex.sort_pairs(keys, pay, (int64_t) nd * npad, npad);
    ex.for_each((int64_t) nd * count, SpIndirections{pay, count, npad, out_ii});
    ex.for_each(shape.n_offsets, SpOffsets{shape, keys, count, npad, out_io});
    sp_check_errors(ex, err);
    res.count = (int32_t) count;
    res.values.resize((size_t) count * sizeof(U));
    res.indices.resize((size_t) count);
    res.offsets.resize((size_t) shape.n_offsets);
    res.indirections.resize((size_t) nd * (size_t) count);
    ex.download(res.values.data(), out_v, count * (int64_t) sizeof(U));
    ex.download(res.indices.data(), out_i, count * 4);
    ex.download(res.offsets.data(), out_io, (int64_t) shape.n_offsets * 4);
    ex.download(res.indirections.data(), out_ii, (int64_t) nd * count * 4);
    ex.sync();
}

// ---- insertSparse (SparseOpsInsert.cpp:26-168) ------------------------------------------------------------------------

template <class U, class Exec>
inline void sp_insert_typed(Exec &ex, const void *oldV, int64_t nold, const int32_t *oldD, const int32_t *oldS,
        const int32_t *oldDo, int nd, const int32_t *oldI, const void *newV, int64_t nnew, int32_t *newI,
        SparseHostResult &res) {
    if (nold < 0 || nnew < 0 || nold > INT_MAX || nnew > INT_MAX || !oldD || !oldS || !oldDo ||
            (nold > 0 && (!oldV || !oldI)) || (nnew > 0 && (!newV || !newI))) {
        throw std::runtime_error("Invalid arguments");
    }
    const SparseShape shape = sp_make_shape(oldD, oldS, oldDo, nd, nold + nnew > 0);
    int32_t *err = ex.template buf<int32_t>(SPB_ERR, SP_ERR_SLOTS);
    ex.zero(err, SP_ERR_SLOTS * 4);
    int32_t *d_old_i = ex.template buf<int32_t>(SPB_OLD_I, nold);
    int32_t *d_new_i = ex.template buf<int32_t>(SPB_NEW_I, nnew);
    U *d_old_v = ex.template buf<U>(SPB_OLD_V, nold);
    U *d_new_v = ex.template buf<U>(SPB_NEW_V, nnew);
    ex.upload(d_old_i, oldI, nold * 4);
    ex.upload(d_new_i, newI, nnew * 4);
    ex.upload(d_old_v, oldV, nold * (int64_t) sizeof(U));
    ex.upload(d_new_v, newV, nnew * (int64_t) sizeof(U));
    // sort the new (index, position) pairs; the caller's newI comes back sorted, as with both reference providers
    // (SparseOpsInsert.cpp:118-126, SparseOps.java:80-84)
    const int64_t npad = sp_next_pow2(nnew);
    uint64_t *keys = ex.template buf<uint64_t>(SPB_KEYS, npad);
    int32_t *pay = ex.template buf<int32_t>(SPB_PAY, npad);
    int32_t *sorted_i = ex.template buf<int32_t>(SPB_SORTED_I, nnew);
    int32_t *perm = ex.template buf<int32_t>(SPB_PERM, nnew);
    int32_t *old_src = ex.template buf<int32_t>(SPB_OLD_SRC, nold);
    ex.for_each(npad, SpNewKeys{d_new_i, nnew, keys, pay});
    ex.sort_pairs(keys, pay, npad, npad);
    ex.for_each(nnew, SpUnpackNew{keys, pay, sorted_i, perm, err});
    ex.for_each(nold, SpIota{old_src});
    ex.download(newI, sorted_i, nnew * 4);
    int32_t h[SP_ERR_SLOTS];
    ex.download(h, err, sizeof(h));
    ex.sync();
    if (h[SP_ERR_DUPLICATE]) {
        throw std::runtime_error("Duplicate values are not allowed");
    }
    sp_merge<U>(ex, shape, err, d_old_i, old_src, (int32_t) nold, d_old_v, sorted_i, perm, (int32_t) nnew, d_new_v, res);
}

template <class Exec>
inline void sp_insert(Exec &ex, int elem_bytes, const void *oldV, int64_t nold, const int32_t *oldD,
        const int32_t *oldS, const int32_t *oldDo, int nd, const int32_t *oldI, const void *newV, int64_t nnew,
        int32_t *newI, SparseHostResult &res) {
    if (elem_bytes == 8) {
        sp_insert_typed<uint64_t>(ex, oldV, nold, oldD, oldS, oldDo, nd, oldI, newV, nnew, newI, res);
    } else if (elem_bytes == 4) {
        sp_insert_typed<uint32_t>(ex, oldV, nold, oldD, oldS, oldDo, nd, oldI, newV, nnew, newI, res);
    } else {
        throw std::runtime_error("Invalid array types");  // NativeArrayKernel.cpp:132-135
    }
}

// ---- sliceSparse (SparseOpsSlice.cpp:26-410) ----------------------------------------------------------------------------

// SparseOpsSlice.cpp:482-497: the indirection ranges of the sliced logical indices have to lie inside the arrays
inline void sp_check_ranges(const std::vector<std::vector<int32_t>> &slices_by_dim, const SparseShape &shape,
        const int32_t *Io, int64_t n) {
    for (int d = 0; d < shape.nd; d++) {
        for (int32_t index : slices_by_dim[(size_t) d]) {
            const int32_t start = Io[shape.dim_offsets[d] + index], end = Io[shape.dim_offsets[d] + index + 1];
            if (!(start >= 0 && start <= end && end >= 0 && end <= n)) {
                throw std::runtime_error("Invalid arguments");
            }
        }
    }
}

template <class U, class Exec>
inline void sp_slice_typed(Exec &ex, const int32_t *slices, int nslices3, const void *srcV, int64_t nsrc,
        const int32_t *srcD, const int32_t *srcS, const int32_t *srcDo, const int32_t *srcI, const int32_t *srcIo,
        const int32_t *srcIi, const void *dstV, int64_t ndst, const int32_t *dstD, const int32_t *dstS,
        const int32_t *dstDo, const int32_t *dstI, const int32_t *dstIo, const int32_t *dstIi, int nd,
        SparseHostResult &res) {
    if (nsrc < 0 || ndst < 0 || nsrc > INT_MAX || ndst > INT_MAX || nslices3 < 0 || nslices3 % 3 != 0 ||
            (nslices3 > 0 && !slices) || !srcD || !srcS || !srcDo || !srcIo || !dstD || !dstS || !dstDo || !dstIo ||
            (nsrc > 0 && (!srcV || !srcI || !srcIi)) || (ndst > 0 && (!dstV || !dstI || !dstIi))) {
        throw std::runtime_error("Invalid arguments");
    }
    const SparseShape src = sp_make_shape(srcD, srcS, srcDo, nd, nsrc > 0);
    const SparseShape dst = sp_make_shape(dstD, dstS, dstDo, nd, nsrc + ndst > 0);
    const int n_slices = nslices3 / 3;
    for (int i = 0; i < nslices3; i += 3) {  // SparseOpsSlice.cpp:146-160
        const int32_t si = slices[i], di = slices[i + 1], dim = slices[i + 2];
        if (!(dim >= 0 && dim < nd)) {
            throw std::runtime_error("Invalid dimension");
        }
        if (!(si >= 0 && si < srcD[dim]) || !(di >= 0 && di < dstD[dim])) {
            throw std::runtime_error("Invalid index");
        }
    }
    // per (dimension, source index): the sorted distinct destination indices it maps to (SparseOpsSlice.cpp:176-243)
    std::vector<std::vector<int32_t>> src_slices((size_t) nd), dst_slices((size_t) nd);
    std::vector<std::vector<int32_t>> lookups((size_t) src.n_offsets);
    for (int i = 0; i < nslices3; i += 3) {
        const int32_t si = slices[i], di = slices[i + 1], dim = slices[i + 2];
        src_slices[(size_t) dim].push_back(si);
        dst_slices[(size_t) dim].push_back(di);
        lookups[(size_t) (src.dim_offsets[dim] + si)].push_back(di);
    }
    auto normalize = [](std::vector<int32_t> &v) {  // SparseOpsSlice.cpp:412-433
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    };
    for (int d = 0; d < nd; d++) {
        normalize(src_slices[(size_t) d]);
        normalize(dst_slices[(size_t) d]);
    }
    std::vector<int32_t> lookup_offsets((size_t) src.n_offsets + 1, 0), lookup_counts((size_t) src.n_offsets + 1, 0);
    std::vector<int32_t> dst_lookups;
    dst_lookups.reserve((size_t) n_slices + 1);
    for (size_t s = 0; s < lookups.size(); s++) {
        normalize(lookups[s]);
        lookup_offsets[s] = (int32_t) dst_lookups.size();
        lookup_counts[s] = (int32_t) lookups[s].size();
        dst_lookups.insert(dst_lookups.end(), lookups[s].begin(), lookups[s].end());
    }
    lookup_offsets[lookups.size()] = (int32_t) dst_lookups.size();
    dst_lookups.push_back(0);  // never empty: keeps the upload and the pointer valid
    std::vector<int32_t> member((size_t) dst.n_offsets + 1, 0);
    for (int d = 0; d < nd; d++) {
        for (int32_t di : dst_slices[(size_t) d]) {
            member[(size_t) (dst.dim_offsets[d] + di)] = 1;
        }
    }
    sp_check_ranges(src_slices, src, srcIo, nsrc);
    sp_check_ranges(dst_slices, dst, dstIo, ndst);

    int32_t *err = ex.template buf<int32_t>(SPB_ERR, SP_ERR_SLOTS);
    ex.zero(err, SP_ERR_SLOTS * 4);
    int32_t *d_src_i = ex.template buf<int32_t>(SPB_NEW_I, nsrc);
    int32_t *d_dst_i = ex.template buf<int32_t>(SPB_OLD_I, ndst);
    U *d_src_v = ex.template buf<U>(SPB_NEW_V, nsrc);
    U *d_dst_v = ex.template buf<U>(SPB_OLD_V, ndst);
    int32_t *d_lo = ex.template buf<int32_t>(SPB_TABLE_A, (int64_t) lookup_offsets.size());
    int32_t *d_lc = ex.template buf<int32_t>(SPB_TABLE_B, (int64_t) lookup_counts.size());
    int32_t *d_dl = ex.template buf<int32_t>(SPB_TABLE_C, (int64_t) dst_lookups.size());
    int32_t *d_member = ex.template buf<int32_t>(SPB_TABLE_D, (int64_t) member.size());
    ex.upload(d_src_i, srcI, nsrc * 4);
    ex.upload(d_dst_i, dstI, ndst * 4);
    ex.upload(d_src_v, srcV, nsrc * (int64_t) sizeof(U));
    ex.upload(d_dst_v, dstV, ndst * (int64_t) sizeof(U));
    ex.upload(d_lo, lookup_offsets.data(), (int64_t) lookup_offsets.size() * 4);
    ex.upload(d_lc, lookup_counts.data(), (int64_t) lookup_counts.size() * 4);
    ex.upload(d_dl, dst_lookups.data(), (int64_t) dst_lookups.size() * 4);
    ex.upload(d_member, member.data(), (int64_t) member.size() * 4);

    // produced (destination index, source element) pairs
    int32_t *maplen = ex.template buf<int32_t>(SPB_MAPLEN, nsrc + 1);
    int64_t *off = ex.template buf<int64_t>(SPB_OFF, nsrc + 1);
    ex.for_each(nsrc, SpMapLen{src, d_src_i, d_lc, maplen, err});
    ex.scan(maplen, nsrc, off);
    const int64_t total = ex.read(off + nsrc);
    sp_check_errors(ex, err);
    if (total > INT_MAX) {
        throw std::runtime_error("Invalid arguments");
    }
    const int64_t npad = sp_next_pow2(total);
    uint64_t *keys = ex.template buf<uint64_t>(SPB_KEYS, npad);
    int32_t *pay = ex.template buf<int32_t>(SPB_PAY, npad);
    int32_t *src_of = ex.template buf<int32_t>(SPB_SRC_OF, total);
    SpExpand expand;
    expand.src = src;
    for (int d = 0; d < kSparseMaxDims; d++) {
        expand.dst_strides[d] = d < nd ? dst.strides[d] : 0;
    }
    expand.src_i = d_src_i;
    expand.lookup_offsets = d_lo;
    expand.lookup_counts = d_lc;
    expand.dst_lookups = d_dl;
    expand.off = off;
    expand.n_src = nsrc;
    expand.total = total;
    expand.keys = keys;
    expand.pay = pay;
    expand.src_of = src_of;
    ex.for_each(npad, expand);
    ex.sort_pairs(keys, pay, npad, npad);
    int32_t *flag = ex.template buf<int32_t>(SPB_FLAG, std::max<int64_t>(total, ndst) + 1);
    int64_t *pos = ex.template buf<int64_t>(SPB_SCAN, std::max<int64_t>(total, ndst) + 1);
    ex.for_each(total, SpFirstFlag{keys, flag});
    ex.scan(flag, total, pos);
    const int64_t n_res = ex.read(pos + total);
    int32_t *res_i = ex.template buf<int32_t>(SPB_RES_I, n_res);
    int32_t *res_src = ex.template buf<int32_t>(SPB_RES_SRC, n_res);
    ex.for_each(total, SpCompactNew{keys, pay, src_of, flag, pos, res_i, res_src});
    // destination elements outside the overwritten region
    ex.for_each(ndst, SpKeepFlag{dst, d_dst_i, d_member, flag, err});
    ex.scan(flag, ndst, pos);
    const int64_t n_kept = ex.read(pos + ndst);
    int32_t *kept_i = ex.template buf<int32_t>(SPB_KEPT_I, n_kept);
    int32_t *kept_src = ex.template buf<int32_t>(SPB_KEPT_SRC, n_kept);
    ex.for_each(ndst, SpCompactOld{d_dst_i, flag, pos, kept_i, kept_src});
    sp_check_errors(ex, err);
    sp_merge<U>(ex, dst, err, kept_i, kept_src, (int32_t) n_kept, d_dst_v, res_i, res_src, (int32_t) n_res, d_src_v, res);
}

template <class Exec>
inline void sp_slice(Exec &ex, int elem_bytes, const int32_t *slices, int nslices3, const void *srcV, int64_t nsrc,
        const int32_t *srcD, const int32_t *srcS, const int32_t *srcDo, const int32_t *srcI, const int32_t *srcIo,
        const int32_t *srcIi, const void *dstV, int64_t ndst, const int32_t *dstD, const int32_t *dstS,
        const int32_t *dstDo, const int32_t *dstI, const int32_t *dstIo, const int32_t *dstIi, int nd,
        SparseHostResult &res) {
    if (elem_bytes == 8) {
        sp_slice_typed<uint64_t>(ex, slices, nslices3, srcV, nsrc, srcD, srcS, srcDo, srcI, srcIo, srcIi, dstV, ndst, dstD,
                dstS, dstDo, dstI, dstIo, dstIi, nd, res);
    } else if (elem_bytes == 4) {
        sp_slice_typed<uint32_t>(ex, slices, nslices3, srcV, nsrc, srcD, srcS, srcDo, srcI, srcIo, srcIi, dstV, ndst, dstD,
                dstS, dstDo, dstI, dstIo, dstIi, nd, res);
    } else {
        throw std::runtime_error("Invalid array types");
    }
}

// ---- the executor's scan, as cooperative-thread-array bodies (emulated under ThreadSanitizer in tests/emu) --------------
//
// Exclusive scan of n int32 into int64, out[n] = total. Three phases: per-chunk sums, one CTA scanning the chunk sums,
// per-chunk scan with the chunk's offset. A chunk is kScanThreads * kScanItems consecutive elements.

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;

template <class Cta>
SSTC_HD inline void scan_chunk_sum(Cta &cta, const int32_t *in, int64_t n, int64_t chunk, int64_t *chunk_sums,
        int64_t *smem) {
    const int tid = cta.tid;
    const int64_t base = chunk * kScanChunk + (int64_t) tid * kScanItems;
    int64_t sum = 0;
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) {
            sum += in[base + k];
        }
    }
    smem[tid] = sum;
    cta.sync();
    for (int w = kScanThreads / 2; w > 0; w >>= 1) {
        if (tid < w) {
            smem[tid] += smem[tid + w];
        }
        cta.sync();
    }
    if (tid == 0) {
        chunk_sums[chunk] = smem[0];
    }
    cta.sync();  // smem is reused by the caller's next chunk
}

// inclusive scan of one value per thread in shared memory; returns this thread's inclusive prefix
template <class Cta>
SSTC_HD inline int64_t scan_cta_inclusive(Cta &cta, int64_t v, int64_t *smem) {
    const int tid = cta.tid;
    smem[tid] = v;
    cta.sync();
    for (int o = 1; o < kScanThreads; o <<= 1) {
        const int64_t t = tid >= o ? smem[tid - o] : 0;
        cta.sync();
        smem[tid] += t;
        cta.sync();
    }
    return smem[tid];
}

// one CTA: chunk_sums[0 .. n_chunks) -> exclusive prefix in place, *total = sum
template <class Cta>
SSTC_HD inline void scan_chunk_offsets(Cta &cta, int64_t *chunk_sums, int64_t n_chunks, int64_t *total, int64_t *smem) {
    const int tid = cta.tid;
    int64_t carry = 0;
    for (int64_t base = 0; base < n_chunks; base += kScanThreads) {
        const int64_t idx = base + tid;
        const int64_t v = idx < n_chunks ? chunk_sums[idx] : 0;
        const int64_t incl = scan_cta_inclusive(cta, v, smem);
        if (idx < n_chunks) {
            chunk_sums[idx] = carry + incl - v;
        }
        carry += smem[kScanThreads - 1];
        cta.sync();  // everybody has read the tile total before the next tile overwrites smem
    }
    if (tid == 0) {
        *total = carry;
    }
}

template <class Cta>
SSTC_HD inline void scan_chunk_write(Cta &cta, const int32_t *in, int64_t n, int64_t chunk, const int64_t *chunk_offsets,
        int64_t *out, int64_t *smem) {
    const int tid = cta.tid;
    const int64_t base = chunk * kScanChunk + (int64_t) tid * kScanItems;
    int64_t sum = 0;
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) {
            sum += in[base + k];
        }
    }
    const int64_t incl = scan_cta_inclusive(cta, sum, smem);
    int64_t run = chunk_offsets[chunk] + incl - sum;
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) {
            out[base + k] = run;
            run += in[base + k];
        }
    }
    cta.sync();  // smem is reused by the caller's next chunk
}

}  // namespace sstc
