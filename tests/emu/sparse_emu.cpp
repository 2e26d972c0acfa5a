// TEST INFRASTRUCTURE ONLY -- runs shared_b200/csrc/sparse_core.cuh (the orchestration and per-element code that
// array_sparse.cu runs on the GPU) with a serial host executor. Exports the two C-ABI entry points under emu_ names so
// that tests/test_emu_sparse.py can compare them with the reference's native SparseOps on the CPU.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>

#include "../../include/sst_cuda.h"
#include "../../shared_b200/csrc/sparse_core.cuh"

namespace {

struct HostExec {
    std::map<int, std::vector<unsigned char>> slots;

    template <class T>
    T *buf(int slot, int64_t n) {
        std::vector<unsigned char> &v = slots[slot];
        // like the device scratch: contents are NOT preserved and not zeroed -- poison them to catch reads of stale data
        v.assign((size_t) (std::max<int64_t>(n, 1) * (int64_t) sizeof(T)), 0xA5);
        return reinterpret_cast<T *>(v.data());
    }
    template <class F>
    void for_each(int64_t n, const F &f) {
        for (int64_t i = n - 1; i >= 0; i--) {  // reverse order: nothing may depend on the order
            f(i);
        }
    }
    void scan(const int32_t *in, int64_t n, int64_t *out) {
        int64_t run = 0;
        for (int64_t i = 0; i < n; i++) {
            out[i] = run;
            run += in[i];
        }
        out[n] = run;
    }
    void sort_pairs(uint64_t *keys, int32_t *pay, int64_t total, int64_t seg) {
        // contract of sort_pairs_segments (array_dimension.cu): every aligned segment ascending by (key, payload)
        if (seg < 2) {
            return;
        }
        std::vector<std::pair<uint64_t, int32_t>> tmp((size_t) seg);
        for (int64_t base = 0; base < total; base += seg) {
            for (int64_t i = 0; i < seg; i++) {
                tmp[(size_t) i] = {keys[base + i], pay[base + i]};
            }
            std::sort(tmp.begin(), tmp.end());
            for (int64_t i = 0; i < seg; i++) {
                keys[base + i] = tmp[(size_t) i].first;
                pay[base + i] = tmp[(size_t) i].second;
            }
        }
    }
    template <class T>
    T read(const T *p) {
        return *p;
    }
    void zero(void *p, int64_t bytes) { memset(p, 0, (size_t) bytes); }
    void upload(void *dst, const void *src, int64_t bytes) {
        if (bytes > 0) {
            memcpy(dst, src, (size_t) bytes);
        }
    }
    void download(void *dst, const void *src, int64_t bytes) {
        if (bytes > 0) {
            memcpy(dst, src, (size_t) bytes);
        }
    }
    void sync() {}
};

thread_local sstc::SparseHostResult t_result;
thread_local std::string t_error;

void publish(int elem_bytes, sstc_sparse_state *out) {
    out->count = t_result.count;
    out->n_offsets = (int32_t) t_result.offsets.size();
    out->n_indirections = (int32_t) t_result.indirections.size();
    out->elem_bytes = elem_bytes;
    out->values = t_result.values.data();
    out->indices = t_result.indices.data();
    out->indirection_offsets = t_result.offsets.data();
    out->indirections = t_result.indirections.data();
}

}  // namespace

extern "C" {

const char *emu_last_error() { return t_error.c_str(); }

int emu_insert_sparse(int elem_bytes, const void *oldV, int64_t nold, const int32_t *oldD, const int32_t *oldS,
        const int32_t *oldDo, int nd, const int32_t *oldI, const void *newV, int64_t nnew, int32_t *newI,
        sstc_sparse_state *out) {
    try {
        memset(out, 0, sizeof(*out));
        HostExec ex;
        sstc::sp_insert(ex, elem_bytes, oldV, nold, oldD, oldS, oldDo, nd, oldI, newV, nnew, newI, t_result);
        publish(elem_bytes, out);
        return 0;
    } catch (const std::exception &e) {
        t_error = e.what();
        return 1;
    }
}

int emu_slice_sparse(int elem_bytes, const int32_t *slices, int nslices3, const void *srcV, int64_t nsrc,
        const int32_t *srcD, const int32_t *srcS, const int32_t *srcDo, const int32_t *srcI, const int32_t *srcIo,
        const int32_t *srcIi, const void *dstV, int64_t ndst, const int32_t *dstD, const int32_t *dstS,
        const int32_t *dstDo, const int32_t *dstI, const int32_t *dstIo, const int32_t *dstIi, int nd,
        sstc_sparse_state *out) {
    try {
        memset(out, 0, sizeof(*out));
        HostExec ex;
        sstc::sp_slice(ex, elem_bytes, slices, nslices3, srcV, nsrc, srcD, srcS, srcDo, srcI, srcIo, srcIi, dstV, ndst,
                dstD, dstS, dstDo, dstI, dstIo, dstIi, nd, t_result);
        publish(elem_bytes, out);
        return 0;
    } catch (const std::exception &e) {
        t_error = e.what();
        return 1;
    }
}

}  // extern "C"
