// rc_fused.cuh -- the last two passes of a plan as ONE kernel whose intermediate never
// leaves the L2 cache.
//
// Two consecutive Stockham passes A (radix RA) and B (radix RB, the last one) only exchange
// data inside "chunks": W adjacent columns rem of pass A, taken for every input row group tB
// of pass B, are produced by RB * W/TA tiles of pass A and consumed by RA * W/TB tiles of
// pass B, RA*RB*W elements (a few MB) in all.  The kernel is a queue of tiles ordered
//     A(chunk 0) .. A(chunk lag-1), A(lag) B(0), A(lag+1) B(1), ...
// one CTA per tile, the queue position taken from the block index (CTAs are dispatched in
// order, so everything a tile waits for has already been issued).  Pass A stores its outputs
// into slot (chunk mod nslot) of a small ring in the layout pass B's TMA boxes want,
// [tB][K_A][W]; a per-chunk counter tells B tiles when their chunk is complete, another one
// tells A tiles when a ring slot may be overwritten.  The ring is re-used every nslot chunks,
// stays resident in the 126 MB L2, and the 8-byte-per-element HBM write + read between the
// two passes disappears.  A spin that exceeds its budget sets an error flag instead of
// hanging the GPU.
#pragma once

#include "rc_exec.cuh"
#include "rc_fft3.cuh"
#include "rc_fft3_inst.cuh"

namespace rc {

struct FusedItem { int role; long long chunk; int idx; };

RC_HD FusedItem fused_decode(const FusedPair& f, long long t) {
    const long long lag_a = (long long)f.lag * f.nA;
    if (t < lag_a) return FusedItem{0, t / f.nA, (int)(t % f.nA)};
    t -= lag_a;
    const long long per = f.nA + f.nB;
    const long long mid = (f.nchunks - f.lag) * per;
    if (t < mid) {
        const long long r = t / per;
        const int pos = (int)(t - r * per);
        if (pos < f.nA) return FusedItem{0, f.lag + r, pos};
        return FusedItem{1, r, pos - f.nA};
    }
    t -= mid;
    return FusedItem{1, f.nchunks - f.lag + t / f.nB, (int)(t % f.nB)};
}

// geometry shared by device code and the host replay
struct FusedTile {
    int b, cc, slot;           // batch entry, chunk within it, ring slot
    int row, w;                // A: tB and tile within the chunk;  B: K_A and tile within the chunk
    long long j0;              // global first column of the tile in its pass
    int valid;                 // columns of the tile that exist (rem < NsA)
};
RC_HD FusedTile fused_tile(const FusedPair& f, const FusedItem& it) {
    FusedTile t;
    t.b = (int)(it.chunk / f.cpb);
    t.cc = (int)(it.chunk - (long long)t.b * f.cpb);
    t.slot = (int)(it.chunk % f.nslot);
    const int T = it.role == 0 ? f.TA : f.TB;
    const int per_row = f.W / T;
    t.row = it.idx / per_row;
    t.w = it.idx - t.row * per_row;
    const long long rem0 = (long long)t.cc * f.W + (long long)t.w * T;
    t.j0 = rem0 + (long long)t.row * f.PA.Ns;
    long long v = f.PA.Ns - rem0;
    t.valid = v < 0 ? 0 : (v > T ? T : (int)v);
    return t;
}

// pass A's outputs of one column pair inside the ring slot
template <class S>
RC_HD V3Out fused_out_ring(const FusedPair& f, const FusedTile& t, int tid) {
    const int cp = tid & (S::CP - 1);
    V3Out o;
    o.act_a = 2 * cp < t.valid;
    o.act_b = 2 * cp + 1 < t.valid;
    o.ns = f.W;
    o.oa = (long long)t.row * S::R * f.W + (long long)t.w * S::T + 2 * cp;
    o.ob = o.oa + 1;
    o.pair = o.act_b;
    return o;
}
// pass B's outputs (the last pass: q = 0) with the columns past NsA masked
template <class S>
RC_HD V3Out fused_out_last(const FusedPair& f, const FusedTile& t, int tid) {
    const int cp = tid & (S::CP - 1);
    V3Out o;
    o.act_a = 2 * cp < t.valid;
    o.act_b = 2 * cp + 1 < t.valid;
    o.ns = f.PB.Ns;
    o.oa = t.j0 + 2 * cp;
    o.ob = o.oa + 1;
    o.pair = f.PB.pair_ok && o.act_b;
    return o;
}

#if defined(__CUDACC__) && !defined(RC_EMULATE)
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fused_wait(const int* cnt, int target, int* err) {
    long long spins = 0;
    while (ld_acquire_gpu(cnt) < target) {
        __nanosleep(100);
        if (++spins > (1LL << 23)) { atomicExch(err, 1); break; }      // ~1 s: give up instead of hanging
    }
}

template <class SA, class SB> struct FusedCfg {
    static constexpr int NT = SA::NT > SB::NT ? SA::NT : SB::NT;
    static constexpr int MINB = SA::MINB < SB::MINB ? SA::MINB : SB::MINB;
    static constexpr int SMEM = SA::SMEM_BYTES > SB::SMEM_BYTES ? SA::SMEM_BYTES : SB::SMEM_BYTES;
};

template <class SA, class SB, int SIGN>
__global__ void __launch_bounds__(FusedCfg<SA, SB>::NT, FusedCfg<SA, SB>::MINB)
v3_fused_ll_kernel(const FusedPair f, const LoadAny ldA, const StoreAny stB, const __grid_constant__ CUtensorMap tmapA,
                   const __grid_constant__ CUtensorMap tmapR) {
    extern __shared__ __align__(128) unsigned char rc_v3_smem[];
    const long long ticket = (long long)blockIdx.x + (long long)blockIdx.y * gridDim.x;
    if (ticket >= f.total()) return;
    const FusedItem it = fused_decode(f, ticket);
    const FusedTile t = fused_tile(f, it);
    const int tid = threadIdx.x;
    float4* tile = (float4*)rc_v3_smem;
    if (t.valid == 0) {                              // tile past the last column of a partial chunk: bookkeeping only
        if (tid == 0) atomicAdd((it.role == 0 ? f.doneA : f.doneB) + it.chunk, 1);
        return;
    }
    if (it.role == 0) {
        float2* tw = (float2*)(rc_v3_smem + (size_t)SA::TILE_F4 * 16);
        uint64_t* bar = (uint64_t*)(tw + SA::R);
        if (tid == 0) v3_issue_tile<SA>(tile, bar, &tmapA, ldA.box_rows, t.j0, t.b);
        V3Tw tws;
        if (tid < SA::NT) {
            v3_load_table<SA, SIGN>(tw, f.PA, tid);
            tws = v3_twiddle_setup<SA, true>(f.PA, t.j0, tid);
        }
        if (tid == 0 && it.chunk >= f.nslot) fused_wait(f.doneB + (it.chunk - f.nslot), f.nB, f.err);   // slot free?
        __syncthreads();
        mbar_wait(bar, 0);
        if (tid < SA::NT) v3_stage0<SA, SIGN, true>(tile, tw, f.PA, V3FromTile<SA::CP>{tile}, t.b, t.j0, tid, tws);
        __syncthreads();
        if constexpr (SA::R1 > 1) {
            if (tid < SA::NT) v3_stage1<SA, SIGN>(tile, tw, tid);
            __syncthreads();
        }
        if (tid < SA::NT) {
            const StoreC64 ring{f.ring + (long long)t.slot * f.slot_elems, 0, 1.0f};
            v3_last_direct<SA, SIGN>(tile, ring, 0, fused_out_ring<SA>(f, t, tid), tid);
        }
        __syncthreads();
        if (tid == 0) {                               // cumulative fence: covers the CTA's stores ordered by the barrier
            __threadfence();
            atomicAdd(f.doneA + it.chunk, 1);
        }
    } else {
        float2* tw = (float2*)(rc_v3_smem + (size_t)SB::TILE_F4 * 16);
        uint64_t* bar = (uint64_t*)(tw + SB::R);
        if (tid == 0) {
            fused_wait(f.doneA + it.chunk, f.nA, f.err);                 // chunk complete?
            asm volatile("fence.proxy.async;" ::: "memory");             // generic-proxy stores -> TMA reads
            mbar_init(bar, 1);
            mbar_expect_tx(bar, (uint32_t)(SB::R * SB::T * sizeof(float2)));
            const int x = 2 * (t.row * f.W + t.w * SB::T);
            for (int r = 0; r < SB::R; r += f.box_rows)
                tma_load_3d(tile + (size_t)r * SB::CP, &tmapR, bar, x, r, t.slot);
        }
        V3Tw tws;
        if (tid < SB::NT) {
            v3_load_table<SB, SIGN>(tw, f.PB, tid);
            tws = v3_twiddle_setup<SB, true>(f.PB, t.j0, tid);
        }
        __syncthreads();
        mbar_wait(bar, 0);
        if (tid == 0) atomicAdd(f.doneB + it.chunk, 1);                  // this tile's input has left the ring
        if (tid < SB::NT) v3_stage0<SB, SIGN, true>(tile, tw, f.PB, V3FromTile<SB::CP>{tile}, t.b, t.j0, tid, tws);
        __syncthreads();
        if constexpr (SB::R1 > 1) {
            if (tid < SB::NT) v3_stage1<SB, SIGN>(tile, tw, tid);
            __syncthreads();
        }
        if (tid < SB::NT) {
            const V3Out o = fused_out_last<SB>(f, t, tid);
            if (stB.kind == kStLmr) v3_last_direct<SB, SIGN>(tile, stB.lmr, t.b, o, tid);
            else if (stB.kind == kStWin) v3_last_direct<SB, SIGN>(tile, stB.win, t.b, o, tid);
            else if (stB.kind == kStAng) v3_last_direct<SB, SIGN>(tile, stB.angle, t.b, o, tid);
            else v3_last_direct<SB, SIGN>(tile, stB.c64, t.b, o, tid);
        }
    }
}
#endif

// Run the fused pair.  ldA: complex64 source of pass A (TMA-described), stB: the plan's final StoreOp.
template <class SA, class SB, int SIGN>
cudaError_t v3_run_fused_ll(const FusedPair& f, const LoadAny& ldA, const StoreAny& stB, cudaStream_t stream) {
#ifdef RC_EMULATE
    (void)stream;
    // host replay: the queue in order (every dependency of a tile precedes it)
    std::vector<float4> smv((size_t)(SA::SMEM_BYTES > SB::SMEM_BYTES ? SA::SMEM_BYTES : SB::SMEM_BYTES) / 16 + 1);
    float4* tile = smv.data();
    for (long long ticket = 0; ticket < f.total(); ticket++) {
        const FusedItem it = fused_decode(f, ticket);
        const FusedTile t = fused_tile(f, it);
        if (t.valid == 0) continue;
        if (it.role == 0) {
            float2* tw = (float2*)(tile + SA::TILE_F4);
            for (int tid = 0; tid < SA::NT; tid++) v3_load_table<SA, SIGN>(tw, f.PA, tid);
            v3_emulate_tma<SA>(tile, ldA.c64, f.PA, t.b, t.j0);
            for (int tid = 0; tid < SA::NT; tid++)
                v3_stage0<SA, SIGN, true>(tile, tw, f.PA, V3FromTile<SA::CP>{tile}, t.b, t.j0, tid, v3_twiddle_setup<SA, true>(f.PA, t.j0, tid));
            if constexpr (SA::R1 > 1) for (int tid = 0; tid < SA::NT; tid++) v3_stage1<SA, SIGN>(tile, tw, tid);
            const StoreC64 ring{f.ring + (long long)t.slot * f.slot_elems, 0, 1.0f};
            for (int tid = 0; tid < SA::NT; tid++) v3_last_direct<SA, SIGN>(tile, ring, 0, fused_out_ring<SA>(f, t, tid), tid);
        } else {
            float2* tw = (float2*)(tile + SB::TILE_F4);
            for (int tid = 0; tid < SB::NT; tid++) v3_load_table<SB, SIGN>(tw, f.PB, tid);
            const float2* slot = f.ring + (long long)t.slot * f.slot_elems;
            float2* tl = (float2*)tile;
            for (int r = 0; r < SB::R; r++)
                for (int c = 0; c < SB::T; c++) {
                    const long long col = (long long)t.row * f.W + (long long)t.w * SB::T + c;
                    tl[r * SB::T + c] = col < (long long)SA::R * f.W ? slot[(long long)r * SA::R * f.W + col] : make_float2(0.f, 0.f);
                }
            for (int tid = 0; tid < SB::NT; tid++)
                v3_stage0<SB, SIGN, true>(tile, tw, f.PB, V3FromTile<SB::CP>{tile}, t.b, t.j0, tid, v3_twiddle_setup<SB, true>(f.PB, t.j0, tid));
            if constexpr (SB::R1 > 1) for (int tid = 0; tid < SB::NT; tid++) v3_stage1<SB, SIGN>(tile, tw, tid);
            for (int tid = 0; tid < SB::NT; tid++) {
                const V3Out o = fused_out_last<SB>(f, t, tid);
                if (stB.kind == kStLmr) v3_last_direct<SB, SIGN>(tile, stB.lmr, t.b, o, tid);
                else if (stB.kind == kStWin) v3_last_direct<SB, SIGN>(tile, stB.win, t.b, o, tid);
                else if (stB.kind == kStAng) v3_last_direct<SB, SIGN>(tile, stB.angle, t.b, o, tid);
                else v3_last_direct<SB, SIGN>(tile, stB.c64, t.b, o, tid);
            }
        }
    }
    return cudaSuccess;
#else
    typedef FusedCfg<SA, SB> Cfg;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(v3_fused_ll_kernel<SA, SB, SIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    // the ring as a tensor: [nslot][RB rows][RA*W columns]
    CUtensorMap tmapR;
    TileSource rs{f.ring, (long long)SA::R * f.W, f.slot_elems, (long long)SA::R * f.W, SB::R, f.nslot};
    if (!tma_source_ok(rs) || !tma_encode_tile_map(&tmapR, rs, f.box_rows, SB::T)) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(f.doneA, 0, sizeof(int) * (size_t)(2 * f.nchunks), stream);
    if (e != cudaSuccess) return e;
    const long long total = (f.total() + 1) / 2 * 2;
    long long gx = total, gy = 1;
    while (gx > 0x40000000LL) { gy *= 2; gx = ((total + gy - 1) / gy + 1) / 2 * 2; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)gx, (unsigned)gy, 1);
    cfg.blockDim = dim3((unsigned)Cfg::NT);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, v3_fused_ll_kernel<SA, SB, SIGN>, f, ldA, stB, ldA.tmap, tmapR);
#endif
}

// schedule types by id, and the (pass A, pass B) schedule pairs the fused kernel is compiled for
template <int ID> struct V3ById;
#define RC_V3_BYID(id, r0, r1, r2, nt, mb, cp, role) \
    template <> struct V3ById<id> { typedef V3Sched<r0, r1, r2, nt, mb, cp> type; };
RC_V3_ALL(RC_V3_BYID)
#undef RC_V3_BYID

}  // namespace rc
