// halo.cuh — slab halo exchange: the peer-memory transport (IPC-mapped arenas, one push kernel with
// the free/ready handshake) and the NCCL send/recv fallback.  Included by fluid.cu inside its
// anonymous namespace, after `struct fluid`, fail(), CU() and check_launch() are defined.
#pragma once

// ---- halo exchange (NCCL point-to-point; all buffers and both neighbours in ONE group) ------------
// For each item: sends my top `n` owned rows up and my bottom `n` owned rows down, receives the
// neighbours' rows into my ghost rows.  `base` is local row 0 of a buffer whose local row 0 is
// global row `off`.  Halo messages are latency-bound (a 4096-float row is 16 KiB), so everything
// that can travel together does.
// `which` names the buffer for the peer-memory path: 0 velocity.read, 1 pressure.read, 2 dye.read,
// 3 divergence (the neighbour's copy of the same buffer is found through its exported offsets and
// the SPMD-identical swap parity).
enum { HB_VELOCITY = 0, HB_PRESSURE = 1, HB_DYE = 2, HB_DIVERGENCE = 3 };
struct HaloItem { void* base; size_t row_bytes; int off, r0, r1, n; int which; };

struct ScanArgs { const float* div; unsigned char* map; int W, row_off, lo0, hi0, lo1, hi1; };
struct PushSeg { const float4* src; float4* dst; unsigned long long n4; };
struct PushArgs {
    PushSeg seg[8];
    int count;
    unsigned seq;
    unsigned* my_flags;        // [0] ready-from-below [1] ready-from-above [2] free-from-below [3] free-from-above
    unsigned* peer_flags[2];   // the same four words in the neighbours' arenas (IPC-mapped); [0] below, [1] above
    int present[2];
    unsigned* counter;         // block-completion counter (local)
    int* err;                  // set when a bounded spin times out (reported as FLUID_ERR_HALO)
    unsigned long long* dbg;   // FLUID_DEBUG_HALO_TIMING: [0] sum wait-free ns [1] sum push ns [2] sum wait-ready ns [3] count [4] t_start
    ScanArgs scan;             // divergence ghost strips to scan for tiny values once they have arrived (W == 0: none)
};

using fk::ld_acquire_sys; using fk::st_release_sys; using fk::global_ns; using fk::spin_until;   // jacobi.cuh

// One halo exchange in ONE kernel, producer and consumer side:
//   (a) tell the neighbours my ghost rows may be overwritten (this kernel is stream-ordered after
//       every kernel of mine that read them), (b) wait until theirs are free, (c) store my boundary
//       rows straight into their ghost rows — dst pointers are the neighbours' arenas mapped through
//       CUDA IPC, so the stores travel over NVLink / NVSwitch — and (d) once the last block is
//       done, release-store "ready = seq" into the neighbours' flag words.
__global__ void __launch_bounds__(256) halo_push_kernel(PushArgs a) {
    if (threadIdx.x == 0) {
        unsigned long long t0 = 0;
        if (blockIdx.x == 0) {
            if (a.dbg) { t0 = global_ns(); a.dbg[4] = t0; }
            for (int side = 0; side < 2; ++side)
                if (a.present[side]) st_release_sys(a.peer_flags[side] + (side == 0 ? 3 : 2), a.seq);
        }
        for (int side = 0; side < 2; ++side)
            if (a.present[side]) spin_until(a.my_flags + (side == 0 ? 2 : 3), a.seq, a.err);
        if (blockIdx.x == 0 && a.dbg) atomicAdd(a.dbg + 0, global_ns() - t0);
    }
    __syncthreads();
    for (int k = 0; k < a.count; ++k) {
        const float4* __restrict__ s = a.seg[k].src;
        float4* __restrict__ d = a.seg[k].dst;
        const unsigned long long n4 = a.seg[k].n4;
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
             i += (unsigned long long)gridDim.x * blockDim.x)
            d[i] = s[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(a.counter, 1u);
        if (done == gridDim.x - 1) {
            *a.counter = 0;
            __threadfence_system();
            for (int side = 0; side < 2; ++side)
                if (a.present[side]) st_release_sys(a.peer_flags[side] + (side == 0 ? 1 : 0), a.seq);
            if (a.dbg) { atomicAdd(a.dbg + 1, global_ns() - a.dbg[4]); atomicAdd(a.dbg + 3, 1ull); }
        }
    }
    // (e) consumer side in the SAME kernel (every block is resident: the grid is at most 2 blocks per SM):
    // acquire the neighbours' "ready" words, then (f) flag tiny values in the divergence strips that arrived.
    if (threadIdx.x == 0) {
        const unsigned long long t0 = (a.dbg && blockIdx.x == 0) ? global_ns() : 0;
        if (a.present[0]) spin_until(a.my_flags + 0, a.seq, a.err);
        if (a.present[1]) spin_until(a.my_flags + 1, a.seq, a.err);
        if (a.dbg && blockIdx.x == 0) atomicAdd(a.dbg + 2, global_ns() - t0);
    }
    __syncthreads();
    if (a.scan.W > 0) {
        const ScanArgs& sc = a.scan;
        const int n0 = max(sc.hi0 - sc.lo0, 0) * sc.W, n1 = max(sc.hi1 - sc.lo1, 0) * sc.W;
        const int mw = fk::tiny_map_w(sc.W);
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n0 + n1; k += gridDim.x * blockDim.x) {
            const int q = k < n0 ? k : k - n0;
            const int j = (k < n0 ? sc.lo0 : sc.lo1) + q / sc.W, i = q % sc.W;
            if (fk::is_tiny_div(sc.div[(size_t)(j - sc.row_off) * sc.W + i])) sc.map[(j / fk::TINY_CH) * mw + i / fk::TINY_CW] = 1;
        }
    }
}

int exchange_p2p(fluid_t* h, const HaloItem* it, int count) {
    PushArgs pa{};
    pa.seq = ++h->p2p_seq;
    pa.my_flags = (unsigned*)(h->arena + h->off_flags);
    pa.counter = pa.my_flags + 8;
    pa.err = h->halo_flag;
    static const bool dbg = getenv("FLUID_DEBUG_HALO_TIMING") != nullptr;
    pa.dbg = dbg ? (unsigned long long*)(h->arena + h->off_flags + 128) : nullptr;
    unsigned long long total4 = 0;
    for (int side = 0; side < 2; ++side) {
        pa.present[side] = h->peer[side].present ? 1 : 0;
        pa.peer_flags[side] = h->peer[side].present ? (unsigned*)(h->peer[side].base + h->peer[side].off_flags) : nullptr;
    }
    for (int k = 0; k < count; ++k) {
        const HaloItem& q = it[k];
        if (q.n <= 0) continue;
        for (int side = 0; side < 2; ++side) {
            const fluid::Peer& P = h->peer[side];
            if (!P.present) continue;
            size_t poff; int proff;
            switch (q.which) {
                case HB_VELOCITY: poff = P.off_v[h->par_v]; proff = P.roff; break;
                case HB_PRESSURE: poff = P.off_p[h->par_p]; proff = P.roff; break;
                case HB_DYE: poff = P.off_dye[h->par_dye]; proff = P.droff; break;
                default: poff = P.off_div; proff = P.roff; break;
            }
            const int g0 = (side == 1) ? q.r1 - q.n : q.r0;          // my owned rows that the neighbour needs
            const char* src = (const char*)q.base + (size_t)(g0 - q.off) * q.row_bytes;
            char* dst = P.base + poff + (size_t)(g0 - proff) * q.row_bytes;
            if (pa.count >= 8) return fail(h, FLUID_ERR_INVALID, "too many halo segments");
            PushSeg& sg = pa.seg[pa.count++];
            sg.src = (const float4*)src; sg.dst = (float4*)dst; sg.n4 = (unsigned long long)q.n * q.row_bytes / 16;
            total4 = std::max(total4, sg.n4);
        }
    }
    const unsigned blocks = (unsigned)std::min<unsigned long long>(std::max<unsigned long long>((total4 + 255) / 256, 1), (unsigned long long)h->sm_count * 2);
    if (h->scan_pending) {          // this exchange carries divergence ghost rows: their tiny-value scan rides along
        pa.scan = ScanArgs{h->scan.div, h->scan.map, h->scan.W, h->scan.row_off, h->scan.lo0, h->scan.hi0, h->scan.lo1, h->scan.hi1};
        h->scan_pending = false;
    }
    halo_push_kernel<<<blocks, 256, 0, h->active>>>(pa);     // push + handshake + wait + scan: one launch per exchange
    int rc = check_launch(h, "halo_push_kernel"); if (rc) return rc;
    h->halo_kernel_launches += 1;
    ++h->halo_groups;
    return FLUID_OK;
}

int exchange_many(fluid_t* h, const HaloItem* it, int count) {
    if (!h->slab()) return FLUID_OK;
    for (int k = 0; k < count; ++k)
        if (it[k].n > it[k].r1 - it[k].r0)
            return fail(h, FLUID_ERR_HALO, "halo of %d rows exceeds the slab height %d (use fewer GPUs or a taller grid)",
                        it[k].n, it[k].r1 - it[k].r0);
    static const bool skip = getenv("FLUID_DEBUG_SKIP_HALO") != nullptr;   // TIMING EXPERIMENTS ONLY: wrong results
    if (skip) return FLUID_OK;
    if (h->p2p) return exchange_p2p(h, it, count);
    ncdl::Api& N = ncdl::api();
    int rc = N.GroupStart();
    for (int k = 0; k < count && !rc; ++k) {
        const HaloItem& q = it[k];
        if (q.n <= 0) continue;
        char* b = static_cast<char*>(q.base);
        auto at = [&](int grow) { return b + (size_t)(grow - q.off) * q.row_bytes; };
        const size_t bytes = (size_t)q.n * q.row_bytes;
        if (h->rank + 1 < h->world) {
            if (!rc) rc = N.Send(at(q.r1 - q.n), bytes, ncdl::ncclInt8, h->rank + 1, h->comm, h->active);
            if (!rc) rc = N.Recv(at(q.r1), bytes, ncdl::ncclInt8, h->rank + 1, h->comm, h->active);
        }
        if (h->rank > 0) {
            if (!rc) rc = N.Send(at(q.r0), bytes, ncdl::ncclInt8, h->rank - 1, h->comm, h->active);
            if (!rc) rc = N.Recv(at(q.r0 - q.n), bytes, ncdl::ncclInt8, h->rank - 1, h->comm, h->active);
        }
    }
    const int rc2 = N.GroupEnd();
    if (rc || rc2) return fail(h, FLUID_ERR_NCCL, "NCCL halo exchange failed: %s", N.GetErrorString(rc ? rc : rc2));
    ++h->halo_groups;
    return FLUID_OK;
}

int exchange_rows(fluid_t* h, int which, void* base, size_t row_bytes, int off, int r0, int r1, int n) {
    HaloItem it{base, row_bytes, off, r0, r1, n, which};
    return exchange_many(h, &it, 1);
}

