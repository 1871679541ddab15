// Sparse workspaces: the [entries][B] block of a workspace as a reserved virtual address range in
// which only the rows a given mode touches are backed by device memory (CUDA virtual memory
// management: cuMemAddressReserve / cuMemCreate / cuMemMap).
//
// Why: in uniform mode without clique beliefs (the streaming pipelines: JT_UNIFORM | JT_NO_BELIEFS)
// the potentials and beliefs of evidence-free cliques are never touched per instance -- on
// config 5 that is 99.98 % of the clique entries, 69 MB of the 80 MB a dense workspace takes per
// instance -- so a dense allocation caps the chunk size at a few hundred instances on a 180 GB
// GPU and the level launches stay small.  The plan, the kernels and every entry offset are
// unchanged: untouched rows simply have no memory behind them (a stray access faults instead of
// silently reading garbage).
//
// The rows are derived from the plan itself: for every task of the launches the mode runs, the
// exact index range of each per-instance operand and output (table maxima), as intervals of
// entries; intervals are scaled by B * itemsize, widened to the allocation granularity and
// merged.  The regions after the block (factor offsets, error counter, uniform workspace) are
// always mapped.  The driver entry points are fetched through cudaGetDriverEntryPoint, so the
// library does not link against libcuda and still loads on a machine without a driver.

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "jt_host.h"

struct jt_sparse_ws {
    CUdeviceptr base = 0;
    size_t va_size = 0, mapped = 0;
    std::vector<std::pair<size_t, size_t>> ranges;             // mapped (offset, size)
    std::vector<CUmemGenericAllocationHandle> handles;
};

namespace {

typedef CUresult (*fn_granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
typedef CUresult (*fn_reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
typedef CUresult (*fn_create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
typedef CUresult (*fn_map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
typedef CUresult (*fn_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
typedef CUresult (*fn_unmap)(CUdeviceptr, size_t);
typedef CUresult (*fn_release)(CUmemGenericAllocationHandle);
typedef CUresult (*fn_free)(CUdeviceptr, size_t);

struct Driver {
    fn_granularity granularity = nullptr;
    fn_reserve reserve = nullptr;
    fn_create create = nullptr;
    fn_map map = nullptr;
    fn_access access = nullptr;
    fn_unmap unmap = nullptr;
    fn_release release = nullptr;
    fn_free free_va = nullptr;
    bool ok = false;
};

template <typename F>
bool entry(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p)
        return false;
    fn = reinterpret_cast<F>(p);
    return true;
}

const Driver& driver() {
    static const Driver d = [] {
        Driver x;
        x.ok = entry("cuMemGetAllocationGranularity", x.granularity) && entry("cuMemAddressReserve", x.reserve) &&
               entry("cuMemCreate", x.create) && entry("cuMemMap", x.map) && entry("cuMemSetAccess", x.access) &&
               entry("cuMemUnmap", x.unmap) && entry("cuMemRelease", x.release) && entry("cuMemAddressFree", x.free_va);
        return x;
    }();
    return d;
}

typedef std::vector<std::pair<long long, long long>> Intervals;   // [begin, end) in entries

long long table_max(const jt_plan* p, int off, int n) {
    int m = 0;
    for (int i = 0; i < n; ++i) m = std::max(m, p->tab[off + i]);
    return m;
}

// Entries of the [entries][B] block that the launches of `flags`' mode read or write per instance.
Intervals touched_entries(const jt_plan* p, int flags) {
    const bool uniform = (flags & JT_UNIFORM) && p->hdr[JT_H_UNI_ENTRIES] > 0;
    const bool beliefs = !(flags & JT_NO_BELIEFS);
    std::vector<int> phases;
    if (uniform) phases = {JT_PHASE_INIT_INSTANCE, JT_PHASE_COLLECT_INSTANCE, JT_PHASE_DIST_PRE_INSTANCE};
    else phases = {JT_PHASE_INIT, JT_PHASE_COLLECT, JT_PHASE_DIST_PRE};
    phases.push_back(beliefs ? JT_PHASE_DIST_MAIN : JT_PHASE_DIST_MAIN_MESSAGES);
    if (!(flags & JT_SKIP_MARGINAL)) phases.push_back(beliefs ? JT_PHASE_MARGINAL : JT_PHASE_MARGINAL_DIRECT);
    Intervals iv;
    auto add = [&](long long lo, long long n) {
        if (n > 0) iv.push_back({lo, lo + n});
    };
    for (const auto& L : p->launches) {
        if (std::find(phases.begin(), phases.end(), L.phase) == phases.end()) continue;
        for (int t = L.begin; t < L.end; ++t) {
            const DTask& k = p->tasks[t];
            const int n_shi = k.n_s / k.n_slo, n_rhi = k.n_r / k.n_rlo;
            const int tf = uniform ? k.flags : 0;
            if (k.kind == JT_KIND_INIT) {
                add(k.out, k.n_s);
                for (int j = k.smsg_begin; j < k.smsg_end; ++j) {
                    const DMsg& m = p->msgs[j];
                    if (m.fid == -2) add(m.off, table_max(p, m.a_hi, n_shi) + table_max(p, m.a_lo, k.n_slo) + 1);
                }
                continue;
            }
            const long long span = table_max(p, k.src_shi, n_shi) + table_max(p, k.src_slo, k.n_slo) +
                                   table_max(p, k.src_rhi, n_rhi) + table_max(p, k.src_rlo, k.n_rlo) + 1;
            if (k.src >= 0 && !(tf & JT_TF_SRC_UNIFORM)) add(k.src, span);
            if (k.beta >= 0 && beliefs) add(k.beta, span);
            if (k.out >= 0 && k.out_space == 0) add(k.out, k.n_s);
            if (k.bel >= 0 && (flags & JT_SEP_BELIEFS)) add(k.bel, k.n_s);
            if (k.own >= 0 && !(tf & JT_TF_OWN_UNIFORM)) add(k.own, k.n_s);
            for (int j = k.rmsg_begin; j < k.smsg_end; ++j) {
                const DMsg& m = p->msgs[j];
                if (uniform && m.uni) continue;
                long long n = table_max(p, m.a_hi, n_shi) + table_max(p, m.a_lo, k.n_slo) + 1;
                if (j < k.rmsg_end) n += table_max(p, m.b_hi, n_rhi) + table_max(p, m.b_lo, k.n_rlo);
                add(m.off, n);
            }
        }
    }
    std::sort(iv.begin(), iv.end());
    Intervals merged;
    for (const auto& x : iv) {
        if (!merged.empty() && x.first <= merged.back().second) merged.back().second = std::max(merged.back().second, x.second);
        else merged.push_back(x);
    }
    return merged;
}

// Byte ranges of the workspace to back with memory, aligned to `gran` and merged.
std::vector<std::pair<size_t, size_t>> byte_ranges(const jt_plan* p, int64_t B, int dtype, int flags, size_t gran,
                                                   size_t work_bytes, size_t total) {
    const size_t w = dtype == JT_F64 ? 8 : 4;
    std::vector<std::pair<size_t, size_t>> r;
    for (const auto& x : touched_entries(p, flags)) {
        const size_t lo = (size_t)x.first * (size_t)B * w / gran * gran;
        const size_t hi = ((size_t)x.second * (size_t)B * w + gran - 1) / gran * gran;
        r.push_back({lo, hi});
    }
    r.push_back({work_bytes / gran * gran, (total + gran - 1) / gran * gran});   // offsets, error counter, uniform workspace
    std::sort(r.begin(), r.end());
    std::vector<std::pair<size_t, size_t>> merged;
    for (const auto& x : r) {
        if (!merged.empty() && x.first <= merged.back().second) merged.back().second = std::max(merged.back().second, x.second);
        else merged.push_back(x);
    }
    return merged;
}

}  // namespace

extern "C" {

int jt_workspace_sparse_rows(const jt_plan* p, int flags, int64_t* intervals, int64_t capacity, int64_t* count) {
    if (!p || !count || capacity < 0 || (capacity > 0 && !intervals)) return jt_fail(JT_ERR_INVALID, "bad argument");
    const Intervals iv = touched_entries(p, flags);
    *count = (int64_t)iv.size();
    for (int64_t i = 0; i < (int64_t)iv.size() && i < capacity; ++i) {
        intervals[2 * i] = iv[i].first;
        intervals[2 * i + 1] = iv[i].second;
    }
    return JT_OK;
}

int jt_workspace_sparse_bytes(const jt_plan* p, int64_t B, int dtype, int flags, size_t* mapped, size_t* dense) {
    if (!p || B <= 0 || (dtype != JT_F32 && dtype != JT_F64) || !mapped) return jt_fail(JT_ERR_INVALID, "bad argument");
    size_t total = 0;
    int rc = jt_workspace_bytes(p, B, dtype, &total);
    if (rc != JT_OK) return rc;
    int64_t lay[4];
    rc = jt_workspace_layout(p, B, dtype, lay);
    if (rc != JT_OK) return rc;
    const size_t gran = 2u << 20;     // the usual minimum granularity; the exact one is used at creation
    size_t sum = 0;
    for (const auto& x : byte_ranges(p, B, dtype, flags, gran, (size_t)lay[0], total)) sum += x.second - x.first;
    *mapped = sum;
    if (dense) *dense = total;
    return JT_OK;
}

int jt_workspace_sparse_create(const jt_plan* p, int64_t B, int dtype, int flags, jt_sparse_ws** out) {
    if (!p || !out || B <= 0 || (dtype != JT_F32 && dtype != JT_F64)) return jt_fail(JT_ERR_INVALID, "bad argument");
    *out = nullptr;
    const Driver& d = driver();
    if (!d.ok) return jt_fail(JT_ERR_CUDA, "CUDA virtual memory management is not available");
    int dev = 0;
    JT_CUDA(cudaGetDevice(&dev));
    JT_CUDA(cudaFree(nullptr));       // make sure the primary context exists
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    size_t gran = 0;
    if (d.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || !gran)
        return jt_fail(JT_ERR_CUDA, "cuMemGetAllocationGranularity failed");
    size_t total = 0;
    int rc = jt_workspace_bytes(p, B, dtype, &total);
    if (rc != JT_OK) return rc;
    int64_t lay[4];
    rc = jt_workspace_layout(p, B, dtype, lay);
    if (rc != JT_OK) return rc;
    jt_sparse_ws* ws = new (std::nothrow) jt_sparse_ws;
    if (!ws) return jt_fail(JT_ERR_NOMEM, "out of host memory");
    ws->va_size = (total + gran - 1) / gran * gran;
    if (d.reserve(&ws->base, ws->va_size, gran, 0, 0) != CUDA_SUCCESS) {
        delete ws;
        return jt_fail(JT_ERR_CUDA, "cuMemAddressReserve of %zu bytes failed", total);
    }
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    for (const auto& x : byte_ranges(p, B, dtype, flags, gran, (size_t)lay[0], total)) {
        const size_t size = std::min(x.second, ws->va_size) - x.first;
        CUmemGenericAllocationHandle h;
        if (d.create(&h, size, &prop, 0) != CUDA_SUCCESS) {
            jt_workspace_sparse_destroy(ws);
            return jt_fail(JT_ERR_NOMEM, "cuMemCreate of %zu bytes failed (out of device memory?)", size);
        }
        ws->handles.push_back(h);
        if (d.map(ws->base + x.first, size, 0, h, 0) != CUDA_SUCCESS ||
            d.access(ws->base + x.first, size, &acc, 1) != CUDA_SUCCESS) {
            jt_workspace_sparse_destroy(ws);
            return jt_fail(JT_ERR_CUDA, "cuMemMap / cuMemSetAccess failed");
        }
        ws->ranges.push_back({x.first, size});
        ws->mapped += size;
    }
    // the error counter starts at zero, like a freshly prepared dense workspace
    cudaError_t e = cudaMemset(reinterpret_cast<void*>(ws->base + (size_t)lay[1]), 0, 256);
    if (e != cudaSuccess) {
        jt_workspace_sparse_destroy(ws);
        return jt_fail(JT_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    }
    *out = ws;
    return JT_OK;
}

void* jt_workspace_sparse_ptr(const jt_sparse_ws* ws) { return ws ? reinterpret_cast<void*>(ws->base) : nullptr; }

size_t jt_workspace_sparse_mapped(const jt_sparse_ws* ws) { return ws ? ws->mapped : 0; }

void jt_workspace_sparse_destroy(jt_sparse_ws* ws) {
    if (!ws) return;
    const Driver& d = driver();
    if (d.ok) {
        cudaDeviceSynchronize();      // unmapping is not stream-ordered: no kernel may still use the rows
        for (const auto& r : ws->ranges) d.unmap(ws->base + r.first, r.second);
        for (CUmemGenericAllocationHandle h : ws->handles) d.release(h);
        if (ws->base) d.free_va(ws->base, ws->va_size);
    }
    delete ws;
}

}  // extern "C"
