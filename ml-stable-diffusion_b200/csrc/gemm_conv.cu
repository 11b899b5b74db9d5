// b200sd -- tcgen05 GEMM / im2col-free implicit-GEMM 3x3 convolution for sm_100a.
//
// One persistent, warp-specialised kernel:
//   warp 0      TMA producer  (cp.async.bulk.tensor 2D/4D boxes, SWIZZLE_128B, mbarrier tx)
//   warp 1      MMA issuer    (one elected lane; tcgen05.mma kind::f16 M=128, N=block_n, K=16;
//                              accumulators double-buffered in TMEM; tcgen05.commit -> mbarriers)
//   warps 2..9  epilogue      (tcgen05.ld 32x32b; two warps per TMEM lane quarter split the columns; bias / time-embedding / GEGLU / residual; fp16|fp32
//                              stores or fp32 split-K partials)
//
// Replaces every nn.Conv2d of the reference UNet / VAE decoder (reference
// python_coreml_stable_diffusion/unet.py:74-84, 435-464, 499-507, 533-551, 601-617, 853, 970).
// The 3x3 convolution never materialises im2col: k-block (tap, 64-channel chunk) is one TMA box
// of the NHWC activation shifted by the tap offset; TMA's out-of-bounds zero fill is the padding.
#include "common.cuh"
#include "../../include/b200sd.h"

#include <algorithm>
#include <cmath>
#include <stdlib.h>

namespace b200sd {

static constexpr int kBM = 128;
static constexpr int kBK = 64;
static constexpr int kAStage = kBM * kBK * 2;  // 16 KiB
static constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
static constexpr int kEpiThreads = 256;
static constexpr int kMaxStages = 8;
static constexpr int kSmemBudget = 220 * 1024;

struct __align__(64) GemmParams {
    CUtensorMap tmA0, tmA1, tmB;
    int mode, M, N, n_store;  // n_store: columns written per row (N or N/2 for GEGLU)
    int C0, Kpt, kc0, kc, taps;
    int kb_total, kb_per_split, splits;
    int m_tiles, n_tiles, block_n, stages;
    int n_img, Hout, Wout, stride, bw_log2, bh_log2, tiles_w, tiles_h;
    int bias_rows, bias_stride, geglu, out_f32;
    int act;         // 0 none, 1 SiLU after bias (generic variant only)
    int cluster;     // split-K on a thread-block cluster: the `splits` CTAs of a tile reduce it through DSMEM
    int pad_lo;      // conv: zero padding before the first row / column (1 = symmetric pad 1; 0 = pad only after: the
                     // VAE encoder's F.pad(x, (0, 1, 0, 1)) + stride-2 conv)
    int two_cta;     // CTA pairs: one tcgen05.mma cta_group::2 computes two M tiles, each CTA stages half of B
    int m_pairs;     // ceil(m_tiles / 2)
    int wgt_tiled;   // B operand pre-tiled: tile (n_tile, kb) starts at row (n_tile * kb_total + kb) * block_n
    int bias_mode;   // 0 none, 1 staged in smem (<= 2 vectors per tile), 2 read from global per chunk
    int res_smem;    // 1: residual tile prefetched into smem with cp.async
    void* out;
    const float* bias;
    const __half* residual;
    float* partial;
};

struct TileCoord {
    int m_tile, n_tile, split;
    int n0, h0, w0;  // conv: output-space origin of the 128-pixel box
};

__device__ __forceinline__ TileCoord decode_work(const GemmParams& p, int work, int cta_rank = 0) {
    TileCoord t;
    if (p.two_cta) {  // work = (pair of M tiles, N tile); the two CTAs of the pair take M tiles 2i and 2i + 1
        t.split = 0;
        t.m_tile = 2 * (work % p.m_pairs) + cta_rank;  // may be == m_tiles for an odd count: an all-padding tile
        t.n_tile = work / p.m_pairs;
    } else if (p.cluster) {  // the CTAs of a cluster (consecutive blockIdx.x) are the k-splits of one tile
        t.split = work % p.splits;
        const int r = work / p.splits;
        t.m_tile = r % p.m_tiles;
        t.n_tile = r / p.m_tiles;
    } else {
        t.m_tile = work % p.m_tiles;
        const int r = work / p.m_tiles;
        t.n_tile = r % p.n_tiles;
        t.split = r / p.n_tiles;
    }
    t.n0 = t.h0 = t.w0 = 0;
    if (p.mode == 1) {
        int tw = t.m_tile % p.tiles_w;
        int r2 = t.m_tile / p.tiles_w;
        int th = r2 % p.tiles_h;
        int tn = r2 / p.tiles_h;
        t.w0 = tw << p.bw_log2;
        t.h0 = th << p.bh_log2;
        t.n0 = tn << (7 - p.bw_log2 - p.bh_log2);
    }
    return t;
}

// k-block range of one split: even floor/ceil distribution on clusters (every split non-empty), fixed stride otherwise
__device__ __forceinline__ void split_range(const GemmParams& p, int split, int& kb0, int& kb1) {
    if (p.cluster) {
        kb0 = split * p.kb_total / p.splits;
        kb1 = (split + 1) * p.kb_total / p.splits;
    } else {
        kb0 = split * p.kb_per_split;
        kb1 = min(kb0 + p.kb_per_split, p.kb_total);
    }
}

// Applies the epilogue to 16 consecutive accumulator columns of one output row and stores them.
// `bias` / `res` are already resolved to this row and column (shared or global memory); null = absent.
// kGeneric = false: compile-time variant for the hot shapes (N % 16 == 0, 16-byte aligned rows): straight-line
// vector code only, which keeps the kernel small enough for the instruction cache of these microsecond kernels.
template <bool kGeneric, bool kGeglu, bool kOutF32, bool kPartial>
__device__ __forceinline__ void epilogue_store16(const GemmParams& p, float (&acc)[16], int out_row, int col0,
                                                 int split, const float* bias, const __half* res) {
    const bool partial = kGeneric ? (p.partial != nullptr) : kPartial;
    const bool geglu = kGeneric ? (p.geglu != 0) : kGeglu;
    const bool out_f32 = kGeneric ? (p.out_f32 != 0) : kOutF32;
    if (partial) {
        float* dst = p.partial + (static_cast<size_t>(split) * p.M + out_row) * p.N + col0;
        if (!kGeneric || (col0 + 16 <= p.N && (p.N & 3) == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (col0 + j < p.N) dst[j] = acc[j];
        }
        return;
    }
    if (bias != nullptr) {
        if (!kGeneric || (col0 + 16 <= p.N && (p.N & 3) == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(bias + j);
                acc[j] += bv.x, acc[j + 1] += bv.y, acc[j + 2] += bv.z, acc[j + 3] += bv.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (col0 + j < p.N) acc[j] += bias[j];
        }
    }
    if (kGeneric && p.act != 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (p.act == 1) acc[j] = silu_f(acc[j]);
            else if (p.act == 2) acc[j] = gelu_erf_f(acc[j]);                              // OpenCLIP MLP ("gelu")
            else acc[j] = __fdividef(acc[j], 1.0f + __expf(-1.702f * acc[j]));             // CLIP "quick_gelu"
        }
    }
    int ocol0 = col0;
    const int nvals = geglu ? 8 : 16;
    if (geglu) {
        // interleaved columns: even = value, odd = gate  (unet.py:616-617: a * gelu(g))
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = acc[2 * j] * gelu_erf_f(acc[2 * j + 1]);
        ocol0 = col0 >> 1;
    }
    const int ld = p.n_store;
    const size_t off = static_cast<size_t>(out_row) * ld + ocol0;
    const bool vec_ok = !kGeneric || ((ocol0 + nvals <= ld) && ((ld & 7) == 0));
    if (res != nullptr) {
        if (vec_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
                if (j < nvals) {
                    const uint4 rv = *reinterpret_cast<const uint4*>(res + j);
                    const __half2* h2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f = __half22float2(h2[q]);
                        acc[j + 2 * q] += f.x;
                        acc[j + 2 * q + 1] += f.y;
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < nvals && ocol0 + j < ld) acc[j] += __half2float(res[j]);
        }
    }
    if (out_f32) {
        float* o = reinterpret_cast<float*>(p.out) + off;
        if (vec_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                if (j < nvals)
                    *reinterpret_cast<float4*>(o + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < nvals && ocol0 + j < ld) o[j] = acc[j];
        }
    } else {
        __half* o = reinterpret_cast<__half*>(p.out) + off;
        if (vec_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
                if (j < nvals) {
                    uint4 pk;
                    pk.x = pack_half2(acc[j], acc[j + 1]);
                    pk.y = pack_half2(acc[j + 2], acc[j + 3]);
                    pk.z = pack_half2(acc[j + 4], acc[j + 5]);
                    pk.w = pack_half2(acc[j + 6], acc[j + 7]);
                    *reinterpret_cast<uint4*>(o + j) = pk;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < nvals && ocol0 + j < ld) o[j] = __float2half_rn(acc[j]);
        }
    }
}

// tile-local row -> (global output row, in-bounds)
__device__ __forceinline__ bool tile_row(const GemmParams& p, const TileCoord& t, int row, int& out_row) {
    if (p.mode == 0) {
        out_row = t.m_tile * kBM + row;
        return out_row < p.M;
    }
    const int dw = row & ((1 << p.bw_log2) - 1);
    const int dh = (row >> p.bw_log2) & ((1 << p.bh_log2) - 1);
    const int dn = row >> (p.bw_log2 + p.bh_log2);
    const int on = t.n0 + dn, oy = t.h0 + dh, ox = t.w0 + dw;
    out_row = (on * p.Hout + oy) * p.Wout + ox;
    return (on < p.n_img) && (oy < p.Hout) && (ox < p.Wout);
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- thread-block cluster primitives (distributed shared memory) ----
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta_rank));
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(ra)
                 : "memory");
    return v;
}

template <bool kGeneric, bool kGeglu, bool kOutF32, bool kPartial, bool kTwoCta = false>
__global__ void __launch_bounds__(kGemmThreads, 1) umma_gemm_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // a CTA of a pair stages only its half of the B tile (block_n / 2 rows); the MMA reads both halves
    const int b_rows = kTwoCta ? (p.block_n >> 1) : p.block_n;
    const int b_stage = b_rows * (kBK * 2);
    const int cta_rank = kTwoCta ? static_cast<int>(cluster_ctarank()) : 0;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + p.stages * kAStage;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + p.stages * b_stage);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full = empty_bar + kMaxStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* bias_s = reinterpret_cast<float*>(tmem_ptr + 4);           // [2][block_n] (16 B aligned)
    __half* res_s = reinterpret_cast<__half*>(bias_s + 2 * 256);      // [128][block_n + 8]
    const int ldr = p.block_n + 8;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.tmA0);
        prefetch_tmap(&p.tmA1);
        prefetch_tmap(&p.tmB);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], kTwoCta ? 2 * kEpiThreads : kEpiThreads);  // pair: both CTAs' epilogues
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (kTwoCta) {
            tmem_alloc_pair(tmem_ptr, 512);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_ptr, 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (kTwoCta) cluster_arrive_wait();  // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // PDL: everything above overlapped the previous kernel's tail; from here on we touch its outputs
    pdl_wait();
    pdl_trigger();

    const int total_work = kTwoCta ? p.m_pairs * p.n_tiles : p.m_tiles * p.n_tiles * p.splits;
    const int work0 = kTwoCta ? (blockIdx.x >> 1) : blockIdx.x;
    const int work_step = kTwoCta ? (gridDim.x >> 1) : gridDim.x;

    if (warp == 0) {
        // ------------------------------- TMA producer -------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = kAStage + b_stage;
            for (int work = work0; work < total_work; work += work_step) {
                const TileCoord t = decode_work(p, work, cta_rank);
                int kb0, kb1;
                split_range(p, t.split, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const int tap = kb / p.kc;
                    const int j = kb - tap * p.kc;
                    const bool src1 = j >= p.kc0;
                    const int c = (src1 ? (j - p.kc0) : j) * kBK;
                    const int wk = tap * p.Kpt + (src1 ? p.C0 : 0) + c;
                    const CUtensorMap* tmA = src1 ? &p.tmA1 : &p.tmA0;
                    void* dst_a = smem_a + stage * kAStage;
                    void* dst_b = smem_b + stage * b_stage;
                    const int brow = p.wgt_tiled ? (t.n_tile * p.kb_total + kb) * p.block_n + cta_rank * b_rows
                                                 : t.n_tile * p.block_n + cta_rank * b_rows;
                    if (kTwoCta) {
                        // both CTAs' loads are credited to the leader's full barrier: it sees 2 x tx_bytes per stage
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * tx_bytes);
                        const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (p.mode == 0) {
                            tma_load_2d_pair(dst_a, tmA, fb, c, t.m_tile * kBM, kEvictNormal);
                        } else {
                            const int r = tap / 3, s3 = tap - 3 * r;
                            tma_load_4d_pair(dst_a, tmA, fb, c, t.w0 * p.stride + s3 - p.pad_lo, t.h0 * p.stride + r - p.pad_lo, t.n0,
                                             kEvictNormal);
                        }
                        tma_load_2d_pair(dst_b, &p.tmB, fb, p.wgt_tiled ? 0 : wk, brow,
                                         p.wgt_tiled ? kEvictFirst : kEvictLast);
                    } else {
                        mbar_expect_tx(&full_bar[stage], tx_bytes);
                        if (p.mode == 0) {
                            tma_load_2d(dst_a, tmA, &full_bar[stage], c, t.m_tile * kBM, kEvictNormal);
                        } else {
                            const int r = tap / 3, s3 = tap - 3 * r;
                            tma_load_4d(dst_a, tmA, &full_bar[stage], c, t.w0 * p.stride + s3 - p.pad_lo,
                                        t.h0 * p.stride + r - p.pad_lo, t.n0, kEvictNormal);
                        }
                        tma_load_2d(dst_b, &p.tmB, &full_bar[stage], p.wgt_tiled ? 0 : wk, brow,
                                    p.wgt_tiled ? kEvictFirst : kEvictLast);
                    }
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer ---------------------------------
        // a pair issues M = 256 (128 rows from each CTA) from the leader only; the peer's MMA warp idles
        const uint32_t idesc = make_idesc_f16(kTwoCta ? 2 * kBM : kBM, p.block_n, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        int iter = 0;
        if (!kTwoCta || cta_rank == 0) {
            for (int work = work0; work < total_work; work += work_step, ++iter) {
                const TileCoord t = decode_work(p, work, cta_rank);
                int kb0, kb1;
                split_range(p, t.split, kb0, kb1);
                const int as = iter & 1;
                const uint32_t aphase = (iter >> 1) & 1;
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * 256;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem_a + stage * kAStage), 1024, 0);
                        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + stage * b_stage), 1024, 0);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            // +32 B per K=16 step inside the 128 B swizzle row (start address is in 16 B units)
                            if (kTwoCta)
                                umma_f16_ss_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                            else
                                umma_f16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        if (kTwoCta) {  // release the stage / publish the accumulator in both CTAs
                            umma_commit_pair(&empty_bar[stage], 3);
                            if (kb == kb1 - 1) umma_commit_pair(&tmem_full[as], 3);
                        } else {
                            umma_commit(&empty_bar[stage]);
                            if (kb == kb1 - 1) umma_commit(&tmem_full[as]);
                        }
                    }
                    __syncwarp();
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else {
        // ------------------------------- epilogue -----------------------------------
        const int lane_group = warp & 3;  // TMEM lanes [32*lane_group, +32) are accessible to this warp
        const int half = (warp - 2) >> 2;  // which of the two warps of this lane quarter: it takes every other chunk
        const int row = lane_group * 32 + lane;
        const int tid_e = threadIdx.x - 64;  // 0..255 among the epilogue warps
        int iter = 0;
        for (int work = work0; work < total_work; work += work_step, ++iter) {
            const TileCoord t = decode_work(p, work, cta_rank);
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            const int ncol0 = t.n_tile * p.block_n;
            int out_row;
            const bool valid = tile_row(p, t, row, out_row);
            // ---- operand prefetch, overlapped with this tile's main loop: bias -> smem, residual -> smem ----
            epi_bar_sync();  // everyone is done reading the staging buffers of the previous tile
            int bias_sel = 0;
            if (p.bias_mode == 1) {
                int row0;
                tile_row(p, t, 0, row0);
                const int img0 = p.bias_rows > 0 ? row0 / p.bias_rows : 0;
                const int nvec = p.bias_rows > 0 ? (p.M + p.bias_rows - 1) / p.bias_rows : 1;
                for (int c = tid_e; c < 2 * p.block_n; c += kEpiThreads) {
                    const int which = c >= p.block_n ? 1 : 0;
                    const int cc = c - which * p.block_n;
                    const int col = ncol0 + cc;
                    float v = 0.f;
                    if (col < p.N && img0 + which < nvec && (which == 0 || p.bias_rows > 0))
                        v = p.bias[static_cast<size_t>(img0 + which) * p.bias_stride + col];
                    bias_s[which * p.block_n + cc] = v;
                }
                if (p.bias_rows > 0 && valid) bias_sel = min(1, max(0, out_row / p.bias_rows - img0));
            }
            if (p.res_smem) {
                const int vpr = p.block_n >> 3;  // 16-byte vectors per tile row
                for (int i = tid_e; i < kBM * vpr; i += kEpiThreads) {
                    const int r = i / vpr, cv = i - r * vpr;
                    int orow;
                    if (tile_row(p, t, r, orow) && ncol0 + cv * 8 < p.N)
                        cp_async16(res_s + r * ldr + cv * 8, p.residual + static_cast<size_t>(orow) * p.n_store + ncol0 + cv * 8);
                }
            }
            const int bias_base = (p.bias_mode == 2 && p.bias_rows > 0 && valid) ? (out_row / p.bias_rows) * p.bias_stride : 0;
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            if (p.res_smem) cp_async_wait_all();
            epi_bar_sync();  // staged bias / residual visible to all epilogue threads
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16) + as * 256;
            auto process16 = [&](float (&acc)[16], int c) {  // c: column offset inside the tile
                const float* bptr = nullptr;
                if (p.bias_mode == 1) bptr = bias_s + bias_sel * p.block_n + c;
                else if (kGeneric && p.bias_mode == 2) bptr = p.bias + bias_base + ncol0 + c;
                const __half* rptr = nullptr;
                if (p.res_smem) rptr = res_s + row * ldr + c;
                else if (kGeneric && p.residual != nullptr)
                    rptr = p.residual + static_cast<size_t>(out_row) * p.n_store + ((ncol0 + c) >> (p.geglu ? 1 : 0));
                epilogue_store16<kGeneric, kGeglu, kOutF32, kPartial>(p, acc, out_row, ncol0 + c, t.split, bptr, rptr);
            };
            auto process32 = [&](const uint32_t (&v)[32], int c) {
                if (kPartial && p.cluster) {
                    // fp32 accumulators -> this CTA's shared memory (over the drained pipeline stages); row stride
                    // block_n + 4 floats keeps the 32 rows of a warp on distinct banks
                    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(smem) + row * (p.block_n + 4) + c);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    return;
                }
                if (!valid) return;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    if (ncol0 + c + 16 * hh < p.N) {
                        float acc[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[16 * hh + j]);
                        process16(acc, c + 16 * hh);
                    }
                }
            };
            if (!kGeneric || (p.block_n & 31) == 0) {
                // this warp's 32-column chunks: half, half+2, ...; software-pipelined (the TMEM load of the next
                // chunk is in flight while the current one is converted and stored)
                uint32_t va[32], vb[32];
                int c = 32 * half;
                if (c < p.block_n) tmem_ld32(taddr + c, va);
                while (c < p.block_n) {
                    tmem_ld_wait();
                    const int c2 = c + 64;
                    if (c2 < p.block_n) tmem_ld32(taddr + c2, vb);
                    process32(va, c);
                    if (c2 >= p.block_n) break;
                    tmem_ld_wait();
                    const int c3 = c2 + 64;
                    if (c3 < p.block_n) tmem_ld32(taddr + c3, va);
                    process32(vb, c2);
                    c = c3;
                }
            } else if (kGeneric) {
                for (int c = 16 * half; c < p.block_n; c += 32) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c, v);
                    tmem_ld_wait();
                    if (valid && ncol0 + c < p.N) {
                        float acc[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
                        process16(acc, c);
                    }
                }
            }
            tc_fence_before();
            if (kTwoCta) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0));  // the leader's MMA warp waits
            else mbar_arrive(&tmem_empty[as]);
        }
    }

    if (kPartial && p.cluster) {
        // ---- split-K reduction inside the cluster: every CTA has parked its fp32 tile in shared memory; CTA `rank`
        // sums its 1/splits slice of the tile over all peers through DSMEM in rank order (deterministic), applies
        // bias / residual / conversion and stores it.  No workspace round trip, no second launch. ----
        __syncwarp();
        cluster_sync_all();
        const TileCoord t = decode_work(p, blockIdx.x);
        const int ncol0 = t.n_tile * p.block_n;
        const int ldred = p.block_n + 4;
        const int q4 = p.block_n >> 2;
        const int total4 = kBM * q4;
        const int lo = static_cast<int>(static_cast<long long>(total4) * t.split / p.splits);
        const int hi = static_cast<int>(static_cast<long long>(total4) * (t.split + 1) / p.splits);
        const uint32_t red_base = smem_u32(smem);
        for (int idx = lo + static_cast<int>(threadIdx.x); idx < hi; idx += kGemmThreads) {
            const int r = idx / q4, c4 = idx - r * q4;
            const int col = ncol0 + c4 * 4;
            int orow;
            if (!tile_row(p, t, r, orow) || col >= p.N) continue;
            const uint32_t la = red_base + static_cast<uint32_t>(r * ldred + c4 * 4) * 4u;
            float4 acc = ld_dsmem_f4(la, 0);
            for (int sp = 1; sp < p.splits; ++sp) {
                const float4 v = ld_dsmem_f4(la, sp);
                acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
            }
            if (p.bias != nullptr) {
                const float4 b = *reinterpret_cast<const float4*>(
                    p.bias + (p.bias_rows > 0 ? (orow / p.bias_rows) * p.bias_stride : 0) + col);
                acc.x += b.x, acc.y += b.y, acc.z += b.z, acc.w += b.w;
            }
            const size_t off = static_cast<size_t>(orow) * p.N + col;
            if (p.residual != nullptr) {
                const uint2 rr = *reinterpret_cast<const uint2*>(p.residual + off);
                const float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(&rr.x));
                const float2 r1 = __half22float2(*reinterpret_cast<const __half2*>(&rr.y));
                acc.x += r0.x, acc.y += r0.y, acc.z += r1.x, acc.w += r1.y;
            }
            if (p.out_f32) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = acc;
            } else {
                uint2 pk;
                pk.x = pack_half2(acc.x, acc.y);
                pk.y = pack_half2(acc.z, acc.w);
                *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + off) = pk;
            }
        }
        cluster_sync_all();  // nobody leaves (or frees its shared memory) while a peer may still read it
    }

    tc_fence_before();
    __syncthreads();
    if (kTwoCta) cluster_arrive_wait();  // the pair retires together: remote barrier arrivals, peer smem / TMEM reads
    if (warp == 1) {
        tc_fence_after();
        if (kTwoCta) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// Sums split-K partials and applies the same epilogue (bias, residual); no GEGLU.
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N,
                                     const float* __restrict__ bias, int bias_rows, int bias_stride,
                                     const __half* __restrict__ residual, void* __restrict__ out, int out_f32) {
    pdl_wait();
    const size_t total4 = static_cast<size_t>(M) * N / 4;
    const size_t stride = static_cast<size_t>(M) * N;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t e = i * 4;
        float4 acc = *reinterpret_cast<const float4*>(partial + e);
        for (int s = 1; s < splits; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(partial + s * stride + e);
            acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
        const int row = static_cast<int>(e / N);
        const int col = static_cast<int>(e - static_cast<size_t>(row) * N);
        if (bias != nullptr) {
            const float* b = bias + (bias_rows > 0 ? (row / bias_rows) * bias_stride : 0) + col;
            acc.x += b[0], acc.y += b[1], acc.z += b[2], acc.w += b[3];
        }
        if (residual != nullptr) {
            const __half2* r = reinterpret_cast<const __half2*>(residual + e);
            const float2 r0 = __half22float2(r[0]), r1 = __half22float2(r[1]);
            acc.x += r0.x, acc.y += r0.y, acc.z += r1.x, acc.w += r1.y;
        }
        if (out_f32) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + e) = acc;
        } else {
            uint2 pk;
            pk.x = pack_half2(acc.x, acc.y);
            pk.y = pack_half2(acc.z, acc.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + e) = pk;
        }
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct GemmPlan {
    int M, N, Kpt, taps, kc0, kc1, kb_total;
    int m_tiles, n_tiles, block_n, splits, kb_per_split, stages;
    int Hout, Wout, bw, bh, bn_img, tiles_w, tiles_h, tiles_n;
    int bias_mode, res_smem, epi_smem;
    int cluster;  // split-K reduced inside a thread-block cluster of `splits` CTAs (DSMEM) instead of a second kernel
    int two_cta;  // CTA pairs (tcgen05.mma cta_group::2, M = 256): each CTA stages half of the B tile
};

static bool cluster_splitk_enabled() {
    // B200SD_CLUSTER_SPLITK=0: always take the workspace + reduce-kernel path (read per call: tuning scripts flip it)
    const char* e = getenv("B200SD_CLUSTER_SPLITK");
    return !(e && e[0] == '0');
}

static bool two_cta_enabled() {
    // opt-in (B200SD_2CTA=1, read per call: tuning scripts flip it).  Correct on every parity test, but measured
    // 2 % slower end to end than single-CTA tiles on the SD-2.1 UNet (199.2 vs 203.5 iter/s): see profiles/README.md
    const char* e = getenv("B200SD_2CTA");
    return e && e[0] == '1';
}

static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

static int plan_gemm(const b200sd_gemm_args& a, GemmPlan& pl) {
    B200SD_REQUIRE(a.mode == 0 || a.mode == 1, "b200sd_gemm: bad mode %d", a.mode);
    B200SD_REQUIRE(!a.pad_after_only || (a.mode == 1 && a.stride == 2), "b200sd_gemm: pad_after_only is for stride-2 convolutions");
    B200SD_REQUIRE(a.c0 > 0 && a.c0 % 8 == 0 && a.c1 >= 0 && a.c1 % 8 == 0,
                   "b200sd_gemm: channel counts must be positive multiples of 8 (c0=%d c1=%d)", a.c0, a.c1);
    B200SD_REQUIRE(a.n > 0, "b200sd_gemm: n=%d", a.n);
    pl.N = a.n;
    pl.Kpt = a.c0 + a.c1;
    pl.kc0 = (a.c0 + kBK - 1) / kBK;
    pl.kc1 = (a.c1 + kBK - 1) / kBK;
    pl.taps = a.mode == 1 ? 9 : 1;
    pl.kb_total = pl.taps * (pl.kc0 + pl.kc1);
    if (a.mode == 0) {
        B200SD_REQUIRE(a.m > 0, "b200sd_gemm: m=%d", a.m);
        pl.M = a.m;
        pl.m_tiles = (a.m + kBM - 1) / kBM;
        pl.Hout = pl.Wout = pl.bw = pl.bh = pl.bn_img = pl.tiles_w = pl.tiles_h = pl.tiles_n = 1;
    } else {
        B200SD_REQUIRE(a.stride == 1 || a.stride == 2, "b200sd_gemm: stride %d", a.stride);
        B200SD_REQUIRE(a.n_img > 0 && a.h > 0 && a.w > 0, "b200sd_gemm: bad image geometry");
        B200SD_REQUIRE(a.stride == 1 || (a.h % 2 == 0 && a.w % 2 == 0), "b200sd_gemm: stride-2 needs even h, w");
        pl.Hout = a.h / a.stride;
        pl.Wout = a.w / a.stride;
        pl.M = a.n_img * pl.Hout * pl.Wout;
        // pick the 128-pixel box (bn_img x bh x bw, powers of two) with the least padding
        long best = -1;
        for (int bw = 128; bw >= 1; bw >>= 1) {
            if (bw * a.stride > 256) continue;
            for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
                if (bh * a.stride > 256) continue;
                const int bn = 128 / (bw * bh);
                const int tw = (pl.Wout + bw - 1) / bw, th = (pl.Hout + bh - 1) / bh, tn = (a.n_img + bn - 1) / bn;
                const long tiles = static_cast<long>(tw) * th * tn;
                // prefer fewer tiles, then wider rows
                const long score = tiles * 1024 - bw;
                if (best < 0 || score < best) {
                    best = score;
                    pl.bw = bw, pl.bh = bh, pl.bn_img = bn, pl.tiles_w = tw, pl.tiles_h = th, pl.tiles_n = tn;
                }
            }
        }
        pl.m_tiles = pl.tiles_w * pl.tiles_h * pl.tiles_n;
    }
    if (a.geglu) B200SD_REQUIRE(a.n % 16 == 0, "b200sd_gemm: GEGLU needs n %% 16 == 0");
    // ---- tile shape / split-K selection by a small cost model (cycles; constants fitted to B200 runs) ----
    const int sms = num_sms();
    const bool can_split = !a.geglu && a.n % 4 == 0 && a.act == 0;
    auto epi_cycles = [&](int bn) { return 400.0 + (bn / 32.0) * (a.geglu ? 520.0 : 230.0); };
    auto kb_cycles = [&](int bn) { return std::max(2.0 * bn, (kAStage + 128.0 * bn) / 38.0); };
    double best_t = 1e30;
    int best_bn = 0, best_s = 1, best_cluster = 0;
    static const int kBns[] = {256, 224, 192, 160, 128, 96, 64, 32, 16};
    static const int kSplits[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32};
    const bool can_cluster = can_split && cluster_splitk_enabled() && a.n % 16 == 0;
    for (int bn : kBns) {
        if (a.block_n > 0 && bn != a.block_n) continue;
        const int nt = (a.n + bn - 1) / bn;
        if (a.block_n == 0 && bn > 16 && nt * bn > a.n + a.n / 4 + 15) continue;  // > 25 % padding
        const int stages_bn = std::min(kMaxStages, (kSmemBudget - 2 * 256 * 4) / (kAStage + bn * kBK * 2));
        for (int sp : kSplits) {
            if (a.split_k > 0 && sp != a.split_k) continue;
            if (sp > 1 && (!can_split || sp * 2 > pl.kb_total) && a.split_k == 0) continue;
            const int kb = (pl.kb_total + sp - 1) / sp;
            const int se = (pl.kb_total + kb - 1) / kb;
            const double main = kb * kb_cycles(bn);
            for (int cl = 0; cl < 2; ++cl) {
                double t;
                if (cl == 1) {
                    // cluster of sp CTAs per tile: portable sizes only, fp32 tile must fit over the pipeline stages
                    if (!can_cluster || (sp != 2 && sp != 4 && sp != 8) || bn % 32 != 0 || sp > pl.kb_total) continue;
                    if (512L * (bn + 4) > static_cast<long>(stages_bn) * (kAStage + bn * kBK * 2)) continue;
                    const long units = static_cast<long>(pl.m_tiles) * nt * sp;
                    // clusters need sp free SMs of one GPC: count ~10 % of the SMs as stranded
                    const double waves = std::ceil(static_cast<double>(units) / (sms - sms / 10));
                    const double epi = 1500.0 + (bn / 32.0) * 120.0 + 128.0 * bn * 4.0 / 17.0;  // TMEM->smem, 2 syncs, DSMEM
                    t = 5000.0 + waves * (main + epi);
                } else {
                    const long units = static_cast<long>(pl.m_tiles) * nt * se;
                    const double waves = std::ceil(static_cast<double>(units) / sms);
                    t = 5000.0 + main + epi_cycles(bn) + (waves - 1.0) * std::max(main, epi_cycles(bn));
                    if (se > 1) t += 6000.0 + static_cast<double>(se) * pl.M * a.n * 8.0 / 3000.0;
                }
                if (t < best_t) {
                    best_t = t;
                    best_bn = bn;
                    best_s = sp;
                    best_cluster = cl;
                }
            }
        }
    }
    if (best_bn == 0) {  // explicit overrides that the loops above did not enumerate
        best_bn = a.block_n > 0 ? a.block_n : 128;
        best_s = a.split_k > 0 ? a.split_k : 1;
    }
    B200SD_REQUIRE(best_bn % 16 == 0 && best_bn >= 16 && best_bn <= 256, "b200sd_gemm: block_n %d", best_bn);
    pl.block_n = best_bn;
    pl.n_tiles = (a.n + pl.block_n - 1) / pl.block_n;
    int splits = std::max(1, std::min(best_s, pl.kb_total));
    pl.kb_per_split = (pl.kb_total + splits - 1) / splits;
    pl.splits = (pl.kb_total + pl.kb_per_split - 1) / pl.kb_per_split;
    pl.cluster = 0;
    if (best_cluster && splits > 1) {  // even distribution, exactly `splits` non-empty ranges (split_range())
        pl.cluster = 1;
        pl.splits = splits;
    }
    if (pl.splits > 1) {
        B200SD_REQUIRE(!a.geglu, "b200sd_gemm: split-K with GEGLU is not supported");
        B200SD_REQUIRE(a.n % 4 == 0, "b200sd_gemm: split-K needs n %% 4 == 0");
    }
    // ---- epilogue operand staging ----
    pl.bias_mode = 0;
    pl.res_smem = 0;
    if (pl.splits == 1) {
        if (a.bias != nullptr) {
            const bool two_vec_ok = a.bias_rows == 0 || a.bias_rows >= kBM ||
                                    (a.mode == 1 && a.bias_rows == pl.Hout * pl.Wout && pl.bn_img <= 2);
            pl.bias_mode = two_vec_ok ? 1 : 2;
        }
        if (a.residual != nullptr && !a.geglu && a.n % 8 == 0) pl.res_smem = 1;
    }
    // ---- CTA pairs: two M tiles share one B tile (each CTA loads half of it), which halves the L2 -> SM operand
    // traffic per flop of B -- the limiter of these GEMMs (~42 B/clk/SM of L2 bandwidth vs 48 KiB per k-block) ----
    const bool regular = (a.act == 0) && (a.n % 16 == 0) && ((a.geglu ? a.n / 2 : a.n) % 8 == 0) && (pl.block_n % 32 == 0) &&
                         pl.bias_mode != 2 && (a.residual == nullptr || pl.res_smem);
    pl.two_cta = (two_cta_enabled() && pl.splits == 1 && regular && pl.m_tiles >= 2 &&
                  static_cast<long>((pl.m_tiles + 1) / 2) * pl.n_tiles * 2 >= num_sms() / 2) ? 1 : 0;
    const int per_stage = kAStage + (pl.two_cta ? pl.block_n / 2 : pl.block_n) * kBK * 2;
    pl.epi_smem = 2 * 256 * 4 + (pl.res_smem ? kBM * (pl.block_n + 8) * 2 : 0);
    pl.stages = std::max(2, std::min(kMaxStages, (kSmemBudget - pl.epi_smem) / per_stage));
    return 0;
}

static size_t plan_workspace(const GemmPlan& pl) {
    return (pl.splits > 1 && !pl.cluster) ? static_cast<size_t>(pl.splits) * pl.M * pl.N * sizeof(float) : 0;
}

extern void count_launch(int n);

static int launch_gemm(const b200sd_gemm_args& a, cudaStream_t stream) {
    GemmPlan pl;
    if (int rc = plan_gemm(a, pl)) return rc;
    B200SD_REQUIRE(a.a0 && a.wgt && a.out, "b200sd_gemm: null pointer");
    B200SD_REQUIRE(a.c1 == 0 || a.a1, "b200sd_gemm: a1 is null but c1 > 0");
    const size_t ws = plan_workspace(pl);
    B200SD_REQUIRE(ws == 0 || (a.workspace && a.workspace_bytes >= ws),
                   "b200sd_gemm: split-K needs %zu workspace bytes, got %zu", ws, a.workspace_bytes);

    GemmParams p;
    memset(&p, 0, sizeof(p));
    // ---- tensor maps ----
    const uint32_t es1[4] = {1, 1, 1, 1};
    if (a.mode == 0) {
        const uint32_t box[2] = {kBK, kBM};
        {
            const uint64_t dims[2] = {static_cast<uint64_t>(a.c0), static_cast<uint64_t>(a.m)};
            const uint64_t str[1] = {static_cast<uint64_t>(a.c0) * 2};
            if (int rc = encode_tmap_f16(&p.tmA0, a.a0, 2, dims, str, box, es1)) return rc;
        }
        if (a.c1 > 0) {
            const uint64_t dims[2] = {static_cast<uint64_t>(a.c1), static_cast<uint64_t>(a.m)};
            const uint64_t str[1] = {static_cast<uint64_t>(a.c1) * 2};
            if (int rc = encode_tmap_f16(&p.tmA1, a.a1, 2, dims, str, box, es1)) return rc;
        } else {
            p.tmA1 = p.tmA0;
        }
    } else {
        const uint32_t st = static_cast<uint32_t>(a.stride);
        const uint32_t box[4] = {kBK, static_cast<uint32_t>(pl.bw) * st, static_cast<uint32_t>(pl.bh) * st,
                                 static_cast<uint32_t>(pl.bn_img)};
        const uint32_t es[4] = {1, st, st, 1};
        for (int src = 0; src < 2; ++src) {
            const int c = src == 0 ? a.c0 : a.c1;
            if (c == 0) {
                p.tmA1 = p.tmA0;
                continue;
            }
            const uint64_t dims[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(a.w),
                                      static_cast<uint64_t>(a.h), static_cast<uint64_t>(a.n_img)};
            const uint64_t str[3] = {static_cast<uint64_t>(c) * 2, static_cast<uint64_t>(c) * 2 * a.w,
                                     static_cast<uint64_t>(c) * 2 * a.w * a.h};
            if (int rc = encode_tmap_f16(src == 0 ? &p.tmA0 : &p.tmA1, src == 0 ? a.a0 : a.a1, 4, dims, str, box, es))
                return rc;
        }
    }
    if (a.wgt_tiled) {
        B200SD_REQUIRE(a.block_n == pl.block_n, "b200sd_gemm: tiled weights need an explicit block_n");
        const uint64_t rows = static_cast<uint64_t>(pl.n_tiles) * pl.kb_total * pl.block_n;
        const uint64_t dims[2] = {kBK, rows};
        const uint64_t str[1] = {kBK * 2};
        const uint32_t box[2] = {kBK, static_cast<uint32_t>(pl.two_cta ? pl.block_n / 2 : pl.block_n)};
        if (int rc = encode_tmap_f16(&p.tmB, a.wgt, 2, dims, str, box, es1)) return rc;
    } else {
        const uint64_t ktot = static_cast<uint64_t>(pl.taps) * pl.Kpt;
        const uint64_t dims[2] = {ktot, static_cast<uint64_t>(a.n)};
        const uint64_t str[1] = {ktot * 2};
        const uint32_t box[2] = {kBK, static_cast<uint32_t>(pl.two_cta ? pl.block_n / 2 : pl.block_n)};
        if (int rc = encode_tmap_f16(&p.tmB, a.wgt, 2, dims, str, box, es1)) return rc;
    }
    p.mode = a.mode;
    p.M = pl.M;
    p.N = a.n;
    p.n_store = a.geglu ? a.n / 2 : a.n;
    p.C0 = a.c0;
    p.Kpt = pl.Kpt;
    p.kc0 = pl.kc0;
    p.kc = pl.kc0 + pl.kc1;
    p.taps = pl.taps;
    p.kb_total = pl.kb_total;
    p.kb_per_split = pl.kb_per_split;
    p.splits = pl.splits;
    p.m_tiles = pl.m_tiles;
    p.n_tiles = pl.n_tiles;
    p.block_n = pl.block_n;
    p.stages = pl.stages;
    p.n_img = a.n_img;
    p.Hout = pl.Hout;
    p.Wout = pl.Wout;
    p.stride = a.mode == 1 ? a.stride : 1;
    p.bw_log2 = ilog2(pl.bw);
    p.bh_log2 = ilog2(pl.bh);
    p.tiles_w = pl.tiles_w;
    p.tiles_h = pl.tiles_h;
    p.bias_rows = a.bias_rows;
    p.bias_stride = a.bias_stride > 0 ? a.bias_stride : a.n;
    p.geglu = a.geglu;
    p.out_f32 = a.out_f32;
    p.act = a.act;
    p.pad_lo = a.pad_after_only ? 0 : 1;
    p.wgt_tiled = a.wgt_tiled;
    p.bias_mode = pl.bias_mode;
    p.res_smem = pl.res_smem;
    p.out = a.out;
    p.bias = a.bias;
    p.residual = reinterpret_cast<const __half*>(a.residual);
    p.partial = (pl.splits > 1 && !pl.cluster) ? a.workspace : nullptr;
    p.cluster = pl.cluster;
    p.two_cta = pl.two_cta;
    p.m_pairs = (pl.m_tiles + 1) / 2;

    const int smem_bytes = pl.stages * (kAStage + (pl.two_cta ? pl.block_n / 2 : pl.block_n) * kBK * 2) + (2 * kMaxStages + 4) * 8 + 16 + pl.epi_smem + 1024;
    const int total = pl.m_tiles * pl.n_tiles * pl.splits;
    const int grid = std::min(total, num_sms());
    // compile-time epilogue variants for the hot shapes; anything irregular takes the generic kernel
    B200SD_REQUIRE(a.act == 0 || (a.act >= 1 && a.act <= 3 && pl.splits == 1 && !a.geglu), "b200sd_gemm: act=%d unsupported here", a.act);
    const bool regular = (a.act == 0) && (a.n % 16 == 0) && (p.n_store % 8 == 0) && (pl.block_n % 32 == 0) && pl.bias_mode != 2 &&
                         (a.residual == nullptr || pl.res_smem || pl.splits > 1);
    using KernelFn = void (*)(GemmParams);
    KernelFn fn;
    int variant;
    if (!regular) {
        fn = umma_gemm_kernel<true, false, false, false>, variant = 0;
    } else if (pl.splits > 1) {
        fn = umma_gemm_kernel<false, false, false, true>, variant = 1;
    } else if (a.geglu) {
        fn = pl.two_cta ? umma_gemm_kernel<false, true, false, false, true> : umma_gemm_kernel<false, true, false, false>;
        variant = pl.two_cta ? 5 : 2;
    } else if (a.out_f32) {
        fn = pl.two_cta ? umma_gemm_kernel<false, false, true, false, true> : umma_gemm_kernel<false, false, true, false>;
        variant = pl.two_cta ? 6 : 3;
    } else {
        fn = pl.two_cta ? umma_gemm_kernel<false, false, false, false, true> : umma_gemm_kernel<false, false, false, false>;
        variant = pl.two_cta ? 7 : 4;
    }
    B200SD_REQUIRE(!pl.two_cta || variant >= 5, "b200sd_gemm: CTA pairs need a regular single-split epilogue variant");
    if (pl.splits > 1 && !pl.cluster) {
        // the separate reduce kernel applies bias / residual; the partial writer must not
        p.bias = nullptr;
        p.residual = nullptr;
    }
    static bool attr_set[8] = {false, false, false, false, false, false, false, false};
    if (!attr_set[variant]) {
        B200SD_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[variant] = true;
    }
    if (pl.cluster || pl.two_cta) {
        B200SD_REQUIRE(pl.two_cta || variant == 1, "b200sd_gemm: cluster split-K needs the regular epilogue variant");
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        // split-K: one (tile, split) per CTA, the splits of a tile are one cluster;  pairs: persistent clusters of 2
        const int pair_units = ((pl.m_tiles + 1) / 2) * pl.n_tiles;
        cfg.gridDim = pl.two_cta ? dim3(2 * std::min(pair_units, num_sms() / 2)) : dim3(total);
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = pl.two_cta ? 2 : pl.splits;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (pdl_enabled()) {
            at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.numAttrs = 2;
        }
        B200SD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, p));
        B200SD_CHECK_CUDA(cudaGetLastError());
        count_launch(1);
        return 0;
    }
    B200SD_CHECK_CUDA(launch_kernel(fn, dim3(grid), dim3(kGemmThreads), smem_bytes, stream, p));
    B200SD_CHECK_CUDA(cudaGetLastError());
    count_launch(1);
    if (pl.splits > 1) {
        const size_t total4 = static_cast<size_t>(pl.M) * a.n / 4;
        const int rgrid = static_cast<int>(std::min<size_t>((total4 + 255) / 256, static_cast<size_t>(num_sms()) * 8));
        B200SD_CHECK_CUDA(launch_kernel(splitk_reduce_kernel, dim3(rgrid), dim3(256), 0, stream, a.workspace, pl.splits, pl.M, a.n, a.bias, a.bias_rows,
                                                        a.bias_stride > 0 ? a.bias_stride : a.n,
                                                        reinterpret_cast<const __half*>(a.residual), a.out, a.out_f32));
        B200SD_CHECK_CUDA(cudaGetLastError());
        count_launch(1);
    }
    return 0;
}

}  // namespace b200sd

extern "C" int b200sd_gemm(const b200sd_gemm_args* args, void* stream) {
    if (!args) {
        b200sd::set_error("b200sd_gemm: args is null");
        return 2;
    }
    return b200sd::launch_gemm(*args, static_cast<cudaStream_t>(stream));
}

extern "C" int b200sd_gemm_plan(const b200sd_gemm_args* args, int32_t* out4) {
    if (!args || !out4) return 2;
    b200sd::GemmPlan pl;
    if (int rc = b200sd::plan_gemm(*args, pl)) return rc;
    out4[0] = pl.block_n, out4[1] = pl.splits, out4[2] = pl.kb_total, out4[3] = pl.n_tiles;
    return 0;
}

extern "C" int b200sd_gemm_describe_plan(const b200sd_gemm_args* args, char* buf, size_t buf_size) {
    if (!args || !buf || buf_size == 0) return 2;
    b200sd::GemmPlan pl;
    if (int rc = b200sd::plan_gemm(*args, pl)) return rc;
    snprintf(buf, buf_size,
             "M=%d N=%d kb_total=%d m_tiles=%d n_tiles=%d block_n=%d splits=%d kb_per_split=%d stages=%d "
             "box=%dx%dx%d bias_mode=%d res_smem=%d epi_smem=%d cluster=%d two_cta=%d",
             pl.M, pl.N, pl.kb_total, pl.m_tiles, pl.n_tiles, pl.block_n, pl.splits, pl.kb_per_split, pl.stages,
             pl.bn_img, pl.bh, pl.bw, pl.bias_mode, pl.res_smem, pl.epi_smem, pl.cluster, pl.two_cta);
    return 0;
}

extern "C" size_t b200sd_gemm_workspace_bytes(const b200sd_gemm_args* args) {
    if (!args) return 0;
    b200sd::GemmPlan pl;
    if (b200sd::plan_gemm(*args, pl)) return 0;
    return b200sd::plan_workspace(pl);
}
