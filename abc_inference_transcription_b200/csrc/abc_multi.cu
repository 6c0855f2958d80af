// abc_multi.cu -- multi-GPU behind the C ABI (abc_init(device_ids, n_dev); 8e).
//
// The reference's parallel model is "start several Julia processes with different `submit` ids and concatenate the files by
// hand" (wrapper.jl:62-63, README.md:37).  Here particles shard across the GPUs of one box by contiguous ranges of the
// global particle index (Philox is keyed by the global index: any partition gives identical bits), with no data-path
// collective.  NCCL over NVLink is used only at the end of a batch:
//   (i)  all-gather of the per-gene acceptance counts of every rank (G uint64 each; their sum is the counts table of
//        model_probs.jl), and
//   (ii) a gene-range exchange of the accepted tuples: genes are cut into n_ranks contiguous ranges of equal tuple mass,
//        every rank orders its own tuples per gene on its GPU (csrc/abc_accept.cu), sends each range to its owner with
//        ncclSend / ncclRecv (one grouped all-to-all), and the owner merges the n_ranks ordered runs of each of its genes
//        with the same three stable radix passes -- so no rank ever holds or sorts the whole accepted set.
// Two front ends share that core:
//   abc_multi_*  one host process (the Julia host): one context + one host thread per GPU, ncclCommInitAll;
//   abc_comm_*   one process per GPU (torchrun / MPI style): ncclCommInitRank from a unique id the host distributes.
// libnccl.so.2 is resolved with dlopen at first use, so the library loads (and every single-GPU entry point works) without it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>

#include "abc_ctx.h"

// ------------------------------------------------------------------------------------------------ NCCL by dlopen
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
        bool ok = true;
        auto sym = [&](const char* name) { void* p = dlsym(api.handle, name); if (!p) { ok = false; api.why = std::string("missing NCCL symbol ") + name; } return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    });
    return api.handle ? &api : nullptr;
}

#define ABC_NCCL_CHECK(expr)                                                                         \
    do {                                                                                             \
        ncclResult_t _r = (expr);                                                                    \
        if (_r != ncclSuccess) {                                                                     \
            abc_set_error("NCCL error %s at %s:%d: %s", #expr, __FILE__, __LINE__, nccl_api()->GetErrorString(_r)); \
            return ABC_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

// per-context communicator state (owned by abc_ctx::comm)
struct AbcComm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    bool owns = false;
    // exchange work buffers
    DevBuf<unsigned long long> d_cnt_all;     // [nranks][G] local counts of every rank
    DevBuf<long long> d_rx_idx, d_mx_idx;     // received tuples (runs by source), merged lists of my gene range
    DevBuf<double> d_rx_err, d_mx_err;
    DevBuf<int32_t> d_rx_gene;
    DevBuf<long long> d_run_start;            // [nranks * G_r + 1]
    // result of the last exchange
    std::vector<int64_t> bounds;              // nranks + 1 gene bounds
    std::vector<int64_t> offsets;             // G + 1 global offsets
    int64_t my_total = 0;
};

void abc_comm_free(abc_ctx* c) {
    AbcComm* m = (AbcComm*)c->comm;
    if (!m) return;
    if (m->comm && m->owns && nccl_api()) nccl_api()->CommDestroy(m->comm);
    delete m;
    c->comm = nullptr;
}

// ------------------------------------------------------------------------------------------------ gene ranges
// n_ranks contiguous gene ranges of (nearly) equal accepted-tuple mass: bounds[k] = first gene of rank k's range.
// Deterministic in the global counts, so every rank computes the same cut.
extern "C" int abc_gene_ranges(const int64_t* counts, int32_t n_genes, int32_t n_ranks, int64_t* bounds) {
    if (!counts || !bounds || n_genes < 0 || n_ranks < 1) { abc_set_error("abc_gene_ranges: bad arguments"); return ABC_ERR_ARG; }
    int64_t total = 0;
    for (int g = 0; g < n_genes; ++g) total += counts[g];
    bounds[0] = 0;
    int64_t cum = 0;
    int g = 0;
    for (int k = 1; k < n_ranks; ++k) {
        // smallest g with cum(g) >= k * total / n_ranks; without any tuple the genes themselves are split evenly
        if (total == 0) { bounds[k] = (int64_t)n_genes * k / n_ranks; continue; }
        const double target = (double)total * (double)k / (double)n_ranks;
        while (g < n_genes && (double)cum < target) { cum += counts[g]; ++g; }
        bounds[k] = g;
    }
    bounds[n_ranks] = n_genes;
    return ABC_OK;
}

// gene of every received tuple: the receive buffer holds, source after source, the runs of the genes of my range
__global__ void abc_rx_gene_kernel(const long long* __restrict__ run_start, int n_runs, int genes_in_range, int g_lo,
                                   long long total, int32_t* __restrict__ gene) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int lo = 0, hi = n_runs;                 // last run with run_start <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (run_start[mid] <= i) lo = mid; else hi = mid;
    }
    gene[i] = g_lo + (lo % genes_in_range);
}

// The exchange.  On return c->comm holds the ordered lists of this rank's gene range (device), the global offsets and the
// gene bounds.  Enqueued on c->stream; synchronised.
static int comm_exchange(abc_ctx* c) {
    AbcComm* m = (AbcComm*)c->comm;
    const int N = m ? m->nranks : 1, me = m ? m->rank : 0, G = c->G;
    if (!c->has_data) { abc_set_error("abc_set_data has not been called"); return ABC_ERR_STATE; }
    if (!m) { abc_set_error("no communicator: call abc_comm_init_rank or use abc_multi_create"); return ABC_ERR_STATE; }
    NcclApi* nc = nccl_api();
    if (N > 1 && !nc) { abc_set_error("NCCL unavailable"); return ABC_ERR_STATE; }
    int rc;
    // local ordered lists (gene, err, particle) and local offsets
    std::vector<int64_t> loff((size_t)G + 1);
    unsigned long long ltotal = 0;
    if ((rc = abc_build_accepted_lists(c, loff.data(), &ltotal, true)) != ABC_OK) return rc;
    // (i) local counts of every rank
    if ((rc = m->d_cnt_all.ensure((size_t)N * G)) != ABC_OK) return rc;
    if (N > 1) {
        ABC_NCCL_CHECK(nc->AllGather(c->d_counts.p, m->d_cnt_all.p, (size_t)G, ncclUint64, m->comm, c->stream));
    } else {
        ABC_CUDA_CHECK(cudaMemcpyAsync(m->d_cnt_all.p, c->d_counts.p, (size_t)G * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
    }
    std::vector<unsigned long long> hall((size_t)N * G);
    ABC_CUDA_CHECK(cudaMemcpyAsync(hall.data(), m->d_cnt_all.p, hall.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    std::vector<int64_t> gc((size_t)G, 0);
    for (int r = 0; r < N; ++r)
        for (int g = 0; g < G; ++g) gc[g] += (int64_t)hall[(size_t)r * G + g];
    m->offsets.assign((size_t)G + 1, 0);
    for (int g = 0; g < G; ++g) m->offsets[g + 1] = m->offsets[g] + gc[g];
    m->bounds.assign((size_t)N + 1, 0);
    if ((rc = abc_gene_ranges(gc.data(), G, N, m->bounds.data())) != ABC_OK) return rc;
    const int g_lo = (int)m->bounds[me], g_hi = (int)m->bounds[me + 1], Gr = g_hi - g_lo;
    // (ii) what comes from whom: run (source r, gene g) has hall[r][g] tuples
    std::vector<long long> run_start((size_t)N * std::max(Gr, 1) + 1, 0);
    std::vector<int64_t> rx_off((size_t)N + 1, 0);
    {
        long long pos = 0;
        for (int r = 0; r < N; ++r) {
            rx_off[r] = pos;
            for (int g = 0; g < Gr; ++g) { run_start[(size_t)r * Gr + g] = pos; pos += (long long)hall[(size_t)r * G + g_lo + g]; }
        }
        rx_off[N] = pos;
        if (Gr > 0) run_start[(size_t)N * Gr] = pos;
    }
    const int64_t rtotal = rx_off[N];
    m->my_total = rtotal;
    if (rtotal != m->offsets[g_hi] - m->offsets[g_lo]) { abc_set_error("gene-range exchange: count mismatch"); return ABC_ERR_STATE; }
    const size_t rcap = (size_t)std::max<int64_t>(rtotal, 1);
    if ((rc = m->d_rx_idx.ensure(rcap)) != ABC_OK) return rc;
    if ((rc = m->d_rx_err.ensure(rcap)) != ABC_OK) return rc;
    if ((rc = m->d_rx_gene.ensure(rcap)) != ABC_OK) return rc;
    if ((rc = m->d_mx_idx.ensure(rcap)) != ABC_OK) return rc;
    if ((rc = m->d_mx_err.ensure(rcap)) != ABC_OK) return rc;
    // my own run: device-to-device; the others: one grouped all-to-all
    {
        const int64_t s0 = loff[g_lo], cnt = loff[g_hi] - loff[g_lo];
        if (cnt > 0) {
            ABC_CUDA_CHECK(cudaMemcpyAsync(m->d_rx_idx.p + rx_off[me], c->d_as_idx.p + s0, (size_t)cnt * sizeof(long long), cudaMemcpyDeviceToDevice, c->stream));
            ABC_CUDA_CHECK(cudaMemcpyAsync(m->d_rx_err.p + rx_off[me], c->d_as_err.p + s0, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    if (N > 1) {
        ABC_NCCL_CHECK(nc->GroupStart());
        for (int r = 0; r < N; ++r) {
            if (r == me) continue;
            const int64_t s0 = loff[m->bounds[r]], scnt = loff[m->bounds[r + 1]] - s0;
            const int64_t rcnt = rx_off[r + 1] - rx_off[r];
            if (scnt > 0) {
                ABC_NCCL_CHECK(nc->Send(c->d_as_idx.p + s0, (size_t)scnt, ncclInt64, r, m->comm, c->stream));
                ABC_NCCL_CHECK(nc->Send(c->d_as_err.p + s0, (size_t)scnt, ncclFloat64, r, m->comm, c->stream));
            }
            if (rcnt > 0) {
                ABC_NCCL_CHECK(nc->Recv(m->d_rx_idx.p + rx_off[r], (size_t)rcnt, ncclInt64, r, m->comm, c->stream));
                ABC_NCCL_CHECK(nc->Recv(m->d_rx_err.p + rx_off[r], (size_t)rcnt, ncclFloat64, r, m->comm, c->stream));
            }
        }
        ABC_NCCL_CHECK(nc->GroupEnd());
    }
    if (rtotal > 0) {
        if (N == 1) {
            // a single rank: the local lists are the result
            ABC_CUDA_CHECK(cudaMemcpyAsync(m->d_mx_idx.p, m->d_rx_idx.p, (size_t)rtotal * sizeof(long long), cudaMemcpyDeviceToDevice, c->stream));
            ABC_CUDA_CHECK(cudaMemcpyAsync(m->d_mx_err.p, m->d_rx_err.p, (size_t)rtotal * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        } else {
            if ((rc = m->d_run_start.ensure(run_start.size())) != ABC_OK) return rc;
            ABC_CUDA_CHECK(cudaMemcpyAsync(m->d_run_start.p, run_start.data(), run_start.size() * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
            const int threads = 256;
            abc_rx_gene_kernel<<<(unsigned)((rtotal + threads - 1) / threads), threads, 0, c->stream>>>(
                m->d_run_start.p, N * Gr, Gr, g_lo, (long long)rtotal, m->d_rx_gene.p);
            ABC_CUDA_CHECK(cudaGetLastError());
            c->launches++;
            // merge the n_ranks ordered runs of every gene: the three stable radix passes of abc_accept.cu
            const size_t tmp_bytes = abc_accept_sort_temp_bytes((size_t)rtotal);
            for (int l = 0; l < 2; ++l) {
                if ((rc = c->d_as_k64[l].ensure((size_t)rtotal)) != ABC_OK) return rc;
                if ((rc = c->d_as_k32[l].ensure((size_t)rtotal)) != ABC_OK) return rc;
                if ((rc = c->d_as_perm[l].ensure((size_t)rtotal)) != ABC_OK) return rc;
            }
            if ((rc = c->d_as_tmp.ensure(tmp_bytes)) != ABC_OK) return rc;
            unsigned long long* pk64[2] = {c->d_as_k64[0].p, c->d_as_k64[1].p};
            uint32_t* pk32[2] = {c->d_as_k32[0].p, c->d_as_k32[1].p};
            uint32_t* pperm[2] = {c->d_as_perm[0].p, c->d_as_perm[1].p};
            int nl = 0;
            if ((rc = abc_launch_accept_sort(m->d_rx_gene.p, m->d_rx_idx.p, m->d_rx_err.p, (size_t)rtotal, G, pk64, pk32, pperm,
                                             c->d_as_tmp.p, tmp_bytes, m->d_mx_idx.p, m->d_mx_err.p, &nl, c->stream)) != ABC_OK) return rc;
            c->launches += nl;
        }
    }
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return ABC_OK;
}

// ------------------------------------------------------------------------------------------------ one process per GPU
extern "C" int abc_comm_unique_id(void* id, size_t bytes) {
    if (!id || bytes < sizeof(ncclUniqueId)) { abc_set_error("abc_comm_unique_id: need a buffer of >= %zu bytes", sizeof(ncclUniqueId)); return ABC_ERR_ARG; }
    NcclApi* nc = nccl_api();
    if (!nc) { abc_set_error("NCCL unavailable: libnccl.so.2 could not be loaded"); return ABC_ERR_STATE; }
    ncclUniqueId u;
    ABC_NCCL_CHECK(nc->GetUniqueId(&u));
    memset(id, 0, bytes);
    memcpy(id, &u, sizeof(u));
    return ABC_OK;
}

extern "C" int abc_comm_init_rank(abc_ctx_t* c, const void* id, size_t bytes, int32_t n_ranks, int32_t rank) {
    if (!c) { abc_set_error("null context"); return ABC_ERR_ARG; }
    ABC_CUDA_CHECK(cudaSetDevice(c->device));
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) { abc_set_error("abc_comm_init_rank: bad rank %d of %d", rank, n_ranks); return ABC_ERR_ARG; }
    abc_comm_free(c);
    AbcComm* m = new (std::nothrow) AbcComm();
    if (!m) { abc_set_error("out of host memory"); return ABC_ERR_NOMEM; }
    m->nranks = n_ranks; m->rank = rank;
    c->comm = m;
    if (n_ranks == 1) return ABC_OK;
    if (!id || bytes < sizeof(ncclUniqueId)) { abc_set_error("abc_comm_init_rank: unique id missing"); return ABC_ERR_ARG; }
    NcclApi* nc = nccl_api();
    if (!nc) { abc_set_error("NCCL unavailable: libnccl.so.2 could not be loaded"); return ABC_ERR_STATE; }
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ABC_NCCL_CHECK(nc->CommInitRank(&m->comm, n_ranks, u, rank));
    m->owns = true;
    return ABC_OK;
}

extern "C" int abc_comm_rank(abc_ctx_t* c, int32_t* n_ranks, int32_t* rank) {
    if (!c || !c->comm) { abc_set_error("no communicator"); return ABC_ERR_STATE; }
    AbcComm* m = (AbcComm*)c->comm;
    if (n_ranks) *n_ranks = m->nranks;
    if (rank) *rank = m->rank;
    return ABC_OK;
}

// per-gene acceptance counts summed over the ranks (host, G int64) -- the counts table of scripts/model_probs.jl
extern "C" int abc_comm_counts(abc_ctx_t* c, int64_t* counts) {
    if (!c) { abc_set_error("null context"); return ABC_ERR_ARG; }
    ABC_CUDA_CHECK(cudaSetDevice(c->device));
    AbcComm* m = (AbcComm*)c->comm;
    if (!m || !c->has_data || !counts) { abc_set_error("abc_comm_counts: bad state/arguments"); return ABC_ERR_STATE; }
    int rc = sync_ctx(c);
    if (rc != ABC_OK) return rc;
    const int G = c->G;
    if ((rc = m->d_cnt_all.ensure((size_t)std::max(m->nranks, 1) * G)) != ABC_OK) return rc;
    if (m->nranks > 1) {
        NcclApi* nc = nccl_api();
        ABC_NCCL_CHECK(nc->AllReduce(c->d_counts.p, m->d_cnt_all.p, (size_t)G, ncclUint64, ncclSum, m->comm, c->stream));
    } else {
        ABC_CUDA_CHECK(cudaMemcpyAsync(m->d_cnt_all.p, c->d_counts.p, (size_t)G * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
    }
    std::vector<unsigned long long> h((size_t)G);
    ABC_CUDA_CHECK(cudaMemcpyAsync(h.data(), m->d_cnt_all.p, (size_t)G * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (int g = 0; g < G; ++g) counts[g] = (int64_t)h[g];
    return ABC_OK;
}

// Collective over the communicator.  Every rank: offsets[G+1] (global), gene_range[2] = the genes this rank ordered.
// root >= 0: that rank also receives the complete lists in idx / errs (sized offsets[G]); the other ranks may pass NULL.
// root < 0: every rank receives only the part of idx / errs that belongs to its own gene range (at the global positions).
extern "C" int abc_comm_accept_fetch(abc_ctx_t* c, int32_t root, int64_t* offsets, int64_t* idx, double* errs, int64_t* gene_range) {
    if (!c) { abc_set_error("null context"); return ABC_ERR_ARG; }
    ABC_CUDA_CHECK(cudaSetDevice(c->device));
    AbcComm* m = (AbcComm*)c->comm;
    if (!m) { abc_set_error("no communicator: call abc_comm_init_rank first"); return ABC_ERR_STATE; }
    if (!offsets || root >= m->nranks) { abc_set_error("abc_comm_accept_fetch: bad arguments"); return ABC_ERR_ARG; }
    int rc = comm_exchange(c);
    if (rc != ABC_OK) return rc;
    const int N = m->nranks, me = m->rank, G = c->G;
    memcpy(offsets, m->offsets.data(), ((size_t)G + 1) * sizeof(int64_t));
    if (gene_range) { gene_range[0] = m->bounds[me]; gene_range[1] = m->bounds[me + 1]; }
    const int64_t my_pos = m->offsets[m->bounds[me]];
    if (root < 0 || N == 1) {
        if (m->my_total > 0) {
            if (idx) ABC_CUDA_CHECK(cudaMemcpyAsync(idx + my_pos, m->d_mx_idx.p, (size_t)m->my_total * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
            if (errs) ABC_CUDA_CHECK(cudaMemcpyAsync(errs + my_pos, m->d_mx_err.p, (size_t)m->my_total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        }
        ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        return ABC_OK;
    }
    // gather the ranges on the root's device, then one copy to its host arrays
    NcclApi* nc = nccl_api();
    const int64_t total = m->offsets[G];
    DevBuf<long long> d_all_idx;
    DevBuf<double> d_all_err;
    if (me == root) {
        if (!idx || !errs) { abc_set_error("abc_comm_accept_fetch: the root needs idx and errs"); return ABC_ERR_ARG; }
        if ((rc = d_all_idx.ensure((size_t)std::max<int64_t>(total, 1))) != ABC_OK) return rc;
        if ((rc = d_all_err.ensure((size_t)std::max<int64_t>(total, 1))) != ABC_OK) return rc;
        if (m->my_total > 0) {
            ABC_CUDA_CHECK(cudaMemcpyAsync(d_all_idx.p + my_pos, m->d_mx_idx.p, (size_t)m->my_total * sizeof(long long), cudaMemcpyDeviceToDevice, c->stream));
            ABC_CUDA_CHECK(cudaMemcpyAsync(d_all_err.p + my_pos, m->d_mx_err.p, (size_t)m->my_total * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    ABC_NCCL_CHECK(nc->GroupStart());
    if (me == root) {
        for (int r = 0; r < N; ++r) {
            if (r == root) continue;
            const int64_t pos = m->offsets[m->bounds[r]], cnt = m->offsets[m->bounds[r + 1]] - pos;
            if (cnt > 0) {
                ABC_NCCL_CHECK(nc->Recv(d_all_idx.p + pos, (size_t)cnt, ncclInt64, r, m->comm, c->stream));
                ABC_NCCL_CHECK(nc->Recv(d_all_err.p + pos, (size_t)cnt, ncclFloat64, r, m->comm, c->stream));
            }
        }
    } else if (m->my_total > 0) {
        ABC_NCCL_CHECK(nc->Send(m->d_mx_idx.p, (size_t)m->my_total, ncclInt64, root, m->comm, c->stream));
        ABC_NCCL_CHECK(nc->Send(m->d_mx_err.p, (size_t)m->my_total, ncclFloat64, root, m->comm, c->stream));
    }
    ABC_NCCL_CHECK(nc->GroupEnd());
    if (me == root && total > 0) {
        ABC_CUDA_CHECK(cudaMemcpyAsync(idx, d_all_idx.p, (size_t)total * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        ABC_CUDA_CHECK(cudaMemcpyAsync(errs, d_all_err.p, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return ABC_OK;
}

// ------------------------------------------------------------------------------------------------ one process, n GPUs
struct abc_multi {
    int n = 0;
    std::vector<abc_ctx*> ctx;
    std::vector<int> dev;
    std::vector<ncclComm_t> comms;
    std::vector<std::string> err;       // last error of every worker thread
    std::vector<int64_t> lo, hi;        // particle shard of every device in the last simulate call
};

// run fn(i) on one host thread per device; the first failure's code and message are returned to the caller's thread
template <typename F>
static int multi_run(abc_multi* mg, F fn) {
    std::vector<int> rcs((size_t)mg->n, ABC_OK);
    std::vector<std::thread> th;
    th.reserve((size_t)mg->n);
    for (int i = 0; i < mg->n; ++i)
        th.emplace_back([&, i] {
            rcs[i] = fn(i);
            if (rcs[i] != ABC_OK) mg->err[i] = abc_last_error();
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < mg->n; ++i)
        if (rcs[i] != ABC_OK) { abc_set_error("device %d: %s", mg->dev[i], mg->err[i].c_str()); return rcs[i]; }
    return ABC_OK;
}

extern "C" int abc_multi_destroy(abc_multi_t* mg) {
    if (!mg) return ABC_OK;
    for (int i = 0; i < mg->n; ++i) if (mg->ctx[i]) abc_destroy(mg->ctx[i]);       // frees the per-context AbcComm (not the comms)
    NcclApi* nc = nccl_api();
    for (auto cm : mg->comms) if (cm && nc) nc->CommDestroy(cm);
    delete mg;
    return ABC_OK;
}

extern "C" int abc_multi_create(const int32_t* devices, int32_t n_dev, abc_multi_t** out) {
    if (!out) { abc_set_error("abc_multi_create: out is NULL"); return ABC_ERR_ARG; }
    *out = nullptr;
    if (n_dev < 1 || n_dev > 64) { abc_set_error("abc_multi_create: n_dev = %d", n_dev); return ABC_ERR_ARG; }
    abc_multi* mg = new (std::nothrow) abc_multi();
    if (!mg) { abc_set_error("out of host memory"); return ABC_ERR_NOMEM; }
    mg->n = n_dev;
    mg->ctx.assign((size_t)n_dev, nullptr);
    mg->err.assign((size_t)n_dev, "");
    mg->lo.assign((size_t)n_dev, 0); mg->hi.assign((size_t)n_dev, 0);
    for (int i = 0; i < n_dev; ++i) mg->dev.push_back(devices ? devices[i] : i);
    for (int i = 0; i < n_dev; ++i)
        for (int j = 0; j < i; ++j)
            if (mg->dev[i] == mg->dev[j]) { abc_set_error("abc_multi_create: device %d listed twice", mg->dev[i]); delete mg; return ABC_ERR_ARG; }
    int rc = ABC_OK;
    for (int i = 0; i < n_dev && rc == ABC_OK; ++i) rc = abc_create(mg->dev[i], &mg->ctx[i]);
    if (rc == ABC_OK && n_dev > 1) {
        NcclApi* nc = nccl_api();
        if (!nc) { abc_set_error("NCCL unavailable: libnccl.so.2 could not be loaded"); rc = ABC_ERR_STATE; }
        else {
            mg->comms.assign((size_t)n_dev, nullptr);
            ncclResult_t r = nc->CommInitAll(mg->comms.data(), n_dev, mg->dev.data());
            if (r != ncclSuccess) { abc_set_error("ncclCommInitAll failed: %s", nc->GetErrorString(r)); rc = ABC_ERR_CUDA; }
        }
    }
    for (int i = 0; i < n_dev && rc == ABC_OK; ++i) {
        AbcComm* m = new (std::nothrow) AbcComm();
        if (!m) { abc_set_error("out of host memory"); rc = ABC_ERR_NOMEM; break; }
        m->nranks = n_dev; m->rank = i; m->owns = false;
        m->comm = (n_dev > 1) ? mg->comms[i] : nullptr;
        mg->ctx[i]->comm = m;
    }
    if (rc != ABC_OK) { std::string keep = abc_last_error(); abc_multi_destroy(mg); abc_set_error("%s", keep.c_str()); return rc; }
    *out = mg;
    return ABC_OK;
}

extern "C" int abc_multi_n_devices(abc_multi_t* mg) { return mg ? mg->n : -1; }
extern "C" abc_ctx_t* abc_multi_ctx(abc_multi_t* mg, int32_t i) { return (mg && i >= 0 && i < mg->n) ? mg->ctx[i] : nullptr; }

extern "C" int abc_multi_set_design(abc_multi_t* mg, const abc_design_t* d) {
    if (!mg) { abc_set_error("null multi context"); return ABC_ERR_ARG; }
    return multi_run(mg, [&](int i) { return abc_set_design(mg->ctx[i], d); });
}
extern "C" int abc_multi_set_data(abc_multi_t* mg, const double* d, const double* se, int32_t G) {
    if (!mg) { abc_set_error("null multi context"); return ABC_ERR_ARG; }
    return multi_run(mg, [&](int i) { return abc_set_data(mg->ctx[i], d, se, G); });
}
extern "C" int abc_multi_set_option(abc_multi_t* mg, const char* name, int64_t value) {
    if (!mg) { abc_set_error("null multi context"); return ABC_ERR_ARG; }
    for (int i = 0; i < mg->n; ++i) { int rc = abc_set_option(mg->ctx[i], name, value); if (rc != ABC_OK) return rc; }
    return ABC_OK;
}
extern "C" int abc_multi_accept_reset(abc_multi_t* mg) {
    if (!mg) { abc_set_error("null multi context"); return ABC_ERR_ARG; }
    return multi_run(mg, [&](int i) { return abc_accept_reset(mg->ctx[i]); });
}
extern "C" int64_t abc_multi_accept_total(abc_multi_t* mg) {
    if (!mg) return -1;
    int64_t t = 0;
    for (int i = 0; i < mg->n; ++i) { int64_t k = abc_accept_total(mg->ctx[i]); if (k < 0) return -1; t += k; }
    return t;
}

// abc_simulate_score over n_trials particles, sharded by contiguous ranges over the devices; the outputs land at their
// global positions in the caller's arrays (same bits as one device).  counts: per-gene counts summed over the devices.
extern "C" int abc_multi_simulate_score(abc_multi_t* mg, int m, int64_t n, int64_t offset, uint64_t seed, int prior_supplied,
                                        double* theta, double* stats, double eps, int layout, double* err, int64_t* counts,
                                        abc_counters_t* counters) {
    if (!mg) { abc_set_error("null multi context"); return ABC_ERR_ARG; }
    const int P = abc_n_params(m);
    if (P < 0) { abc_set_error("model index m = %d must be in 1..5", m); return ABC_ERR_ARG; }
    const int N = mg->n;
    const int64_t base = n / N, rem = n % N;
    for (int i = 0; i < N; ++i) {
        mg->lo[i] = i * base + std::min<int64_t>(i, rem);
        mg->hi[i] = mg->lo[i] + base + (i < rem ? 1 : 0);
    }
    std::vector<abc_counters_t> cn((size_t)N);
    const int G = mg->ctx[0]->G;
    int rc = multi_run(mg, [&](int i) {
        memset(&cn[i], 0, sizeof(abc_counters_t));
        const int64_t lo = mg->lo[i], nb = mg->hi[i] - lo;
        if (nb <= 0) return (int)ABC_OK;
        double* e = nullptr;
        if (err) e = (layout == ABC_ERR_GENE_MAJOR) ? err + lo : err + lo * (int64_t)G;
        return abc_simulate_score_impl(mg->ctx[i], m, nb, offset + lo, seed, prior_supplied, theta + lo * P, stats + lo * ABC_NSTATS,
                                       eps, layout, e, n, nullptr, &cn[i]);
    });
    if (rc != ABC_OK) return rc;
    if (counts) {
        std::vector<int64_t> part((size_t)N * G);
        rc = multi_run(mg, [&](int i) { return abc_comm_counts(mg->ctx[i], part.data() + (size_t)i * G); });   // NCCL all-reduce
        if (rc != ABC_OK) return rc;
        memcpy(counts, part.data(), (size_t)G * sizeof(int64_t));
    }
    if (counters) {
        memset(counters, 0, sizeof(*counters));
        for (int i = 0; i < N; ++i) {
            counters->n_particles += cn[i].n_particles; counters->n_lineages += cn[i].n_lineages; counters->n_events += cn[i].n_events;
            counters->n_draws += cn[i].n_draws; counters->n_ode_steps += cn[i].n_ode_steps;
            counters->ms_simulate = std::max(counters->ms_simulate, cn[i].ms_simulate);
            counters->ms_stats = std::max(counters->ms_stats, cn[i].ms_stats);
            counters->ms_score = std::max(counters->ms_score, cn[i].ms_score);
        }
    }
    return ABC_OK;
}

// merged per-gene ordered lists of all devices: offsets[G+1], idx / errs sized offsets[G] (nullable).  Every device orders
// the genes of its range and copies them to their global positions in the caller's arrays.
extern "C" int abc_multi_accept_fetch(abc_multi_t* mg, int64_t* offsets, int64_t* idx, double* errs) {
    if (!mg || !offsets) { abc_set_error("abc_multi_accept_fetch: bad arguments"); return ABC_ERR_ARG; }
    const int G = mg->ctx[0]->G;
    std::vector<int64_t> off((size_t)mg->n * ((size_t)G + 1));
    int rc = multi_run(mg, [&](int i) {
        return abc_comm_accept_fetch(mg->ctx[i], -1, off.data() + (size_t)i * (G + 1), idx, errs, nullptr);
    });
    if (rc != ABC_OK) return rc;
    memcpy(offsets, off.data(), ((size_t)G + 1) * sizeof(int64_t));
    return ABC_OK;
}
