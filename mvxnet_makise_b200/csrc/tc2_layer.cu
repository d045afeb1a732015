// Stage 3, CTA-pair variant: tcgen05.mma.cta_group::2 (M = 256 across two SMs of a cluster), persistent, with
// double-buffered TMEM accumulators so that the epilogue of tile i runs under the main loop of tile i+1.
//
// Why a pair: a single CTA running SS-mode TF32 MMAs reads (128 + N) x 32 B of shared memory per instruction; with
// the 3xTF32 operand split that is 96 B/cycle of the 128 B/cycle an SM has, and the accumulators of a 256 x 256 tile
// fill all 512 TMEM columns, so the epilogue (21 % of the single-CTA kernel, measured) cannot overlap. In a pair each
// SM holds 128 rows of A and HALF of the B tile (N/2 rows): shared-memory reads per SM drop to 64 B/cycle, the weight
// stream per SM halves, and a 256-column accumulator leaves room for a second one.
//
// Per CTA (rank r of the pair), BK = 16 fp32 per stage, K-major SWIZZLE_64B operands, 3xTF32 split (see tc_layer.cu):
//   warps 0-7  A producers, two groups of four warps on alternate k-chunks (128 rows of the tile: coalesced loads two
//              of the group's chunks ahead, BN of the producer layer, split, swizzled st.shared, fence.proxy.async,
//              arrive on the local full barrier)
//   warp 8     B producer: one cp.async.bulk per stage of this CTA's pre-packed N/2 x 16 [hi|lo] image
//   warp 9     rank 0: MMA issuer (waits its own and the peer's stage, issues 6 MMAs, commits with multicast to both
//              CTAs' empty barriers);  rank 1: relay (forwards "my stage is full" to rank 0's peer_full barrier)
//   warps 10-13 epilogue from registers: tcgen05.ld -> bias/ReLU -> row stores, butterfly column sums (fp32 in a warp,
//              fp64 beyond), then a (remote) arrive on rank 0's tmem_empty barrier
#include "layers.cuh"
#include "tc_common.cuh"

#include <cstdlib>

namespace mvx {

namespace {

constexpr int kRowsPerCta = 128;
constexpr int kThreads2 = 14 * 32;
constexpr int kAProd = 128;       // threads that produce ONE chunk
constexpr int kAGroups = 2;       // producer groups working on alternate chunks (a chunk's store->fence->arrive chain is
                                  // ~1100 cycles, longer than the 768 tensor cycles a pair spends on it)

template <int BN>
struct Smem2 {
    static constexpr int kAHalf = kRowsPerCta * kBK * 4;   // 8 KB: A_hi (then A_lo)
    static constexpr int kBHalf = (BN / 2) * kBK * 4;      // this CTA's half of B_hi (then B_lo)
    static constexpr int kStage = 2 * kAHalf + 2 * kBHalf;
    static constexpr int kStages = (192 * 1024) / kStage;  // 6 stages at BN = 256, 8 at BN = 128
    static constexpr int kTiles = kStages * kStage;
    static constexpr int kMean = kTiles;
    static constexpr int kRstd = kMean + 768 * 4;
    static constexpr int kPart = kRstd + 768 * 4;           // [4 warps][BN][2] fp64 column partials
    static constexpr int kBars = kPart + 4 * BN * 2 * 8;    // full[S], peer_full[S], empty[S], accum_full[2], tmem_empty[2]
    static constexpr int kTmemPtr = kBars + 8 * (3 * kStages + 4);
    static constexpr int kTotal = kTmemPtr + 16 + 1024;
};

// W^T (Cin, Cout) -> per (column tile, k chunk, CTA rank) the shared-memory image [hi | lo] of that rank's N/2 rows
template <int BN>
__global__ void __launch_bounds__(256) pack_weights2_kernel(const float *__restrict__ Wt, int Cin, int Cout, float *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Cin * Cout) return;
    const int k = e / Cout, n = e - k * Cout;
    const int ct = n / BN, nl = n - ct * BN, r = nl / (BN / 2), nr = nl - r * (BN / 2), kc = k / kBK, kl = k - kc * kBK;
    float hi, lo;
    split_tf32(Wt[e], hi, lo);
    constexpr int kHalfElems = (BN / 2) * kBK;
    const size_t blob = (((size_t)ct * (Cin / kBK) + kc) * 2 + r) * (2 * kHalfElems);
    const uint32_t off = sw64_offset(nr, kl >> 2) / 4 + (kl & 3);
    out[blob + off] = hi;
    out[blob + kHalfElems + off] = lo;
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
tc2_layer_kernel(LayerArgs a, const float *__restrict__ wpack, int F, int row_tiles, int col_tiles) {
    using S = Smem2<BN>;
    constexpr int kStages = S::kStages;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    float *s_mean = reinterpret_cast<float *>(smem + S::kMean);
    float *s_rstd = reinterpret_cast<float *>(smem + S::kRstd);
    double *s_part = reinterpret_cast<double *>(smem + S::kPart);
    const uint32_t bars = sbase + S::kBars;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto peer_full_bar = [&](int s) { return bars + 8u * (kStages + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
    auto accum_bar = [&](int b) { return bars + 8u * (3 * kStages + b); };
    auto tmem_empty_bar = [&](int b) { return bars + 8u * (3 * kStages + 2 + b); };
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(smem + S::kTmemPtr);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int nk = a.Cin / kBK;
    const int total = F * row_tiles * col_tiles;
    const int n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;

    auto decode = [&](int t, int &f, int &ct, long long &row0, long long &n_rows) -> bool {
        ct = t % col_tiles;
        const int rt = (t / col_tiles) % row_tiles;
        f = t / (col_tiles * row_tiles);
        n_rows = a.rows_fixed;
        if (a.counts) {
            const int N = a.counts[f * 4 + 0], K = a.counts[f * 4 + 1];
            n_rows = a.rows_mode == 1 ? K + 1 : (a.rows_mode == 2 ? K + N : a.rows_fixed);
        }
        row0 = (long long)rt * (2 * kRowsPerCta);
        return row0 < n_rows;
    };

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), kAProd + 1);
            mbar_init(peer_full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accum_bar(b), 1);
            mbar_init(tmem_empty_bar(b), 8);  // 4 epilogue warps of each CTA of the pair
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {  // same logical warp in both CTAs
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + S::kTmemPtr), "r"(2 * BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const bool mma_only = (a.dbg & 8) != 0;
    if (mma_only && warp != 9) {
        // experiment: nobody but the MMA issuer works
    } else if (warp < 4 * kAGroups) {
        // ================= A producers: this CTA's 128 rows of the pair tile ======================================
        const int grp = warp >> 2, gt = tid & (kAProd - 1);
        const int c = gt & 3, rsub = gt >> 2;  // rows rsub + 32*i
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int gbase = 0, cur_f = -1;               // ring position of the current tile's chunk 0
        for (int t = cluster_id; t < total; t += n_clusters) {
            int f, ct;
            long long row0, n_rows;
            if (!decode(t, f, ct, row0, n_rows)) continue;
            row0 += (long long)rank * kRowsPerCta;
            if (a.in_stats && f != cur_f) {
                named_bar_sync(1, kAProd * kAGroups);
                const double Rstat = a.counts ? (double)a.counts[f * 4 + 0] * (double)a.T : (double)a.rows_fixed;
                for (int cc = tid; cc < a.Cin; cc += kAProd * kAGroups) {
                    const double *st = a.in_stats + ((size_t)f * a.Cin + cc) * 2;
                    const double m = st[0] / Rstat;
                    double var = st[1] / Rstat - m * m;
                    var = var < 0.0 ? 0.0 : var;
                    s_mean[cc] = (float)m;
                    s_rstd[cc] = (float)(1.0 / sqrt(var + a.eps));
                }
                named_bar_sync(1, kAProd * kAGroups);
            }
            cur_f = f;
            const float *Xf = a.X + ((size_t)f * a.rowcap + row0) * a.ldx + c * 4;
            bool valid[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) valid[i] = row0 + rsub + 32 * i < n_rows;
            auto load_chunk = [&](float4 (&buf)[4], int kc) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    buf[i] = (valid[i] && !(a.dbg & 4)) ? __ldg(reinterpret_cast<const float4 *>(Xf + (size_t)(rsub + 32 * i) * a.ldx + kc * kBK)) : z4;
            };
            auto produce = [&](float4 (&buf)[4], int kc) {
                float4 cur[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) cur[i] = buf[i];
                if (a.in_stats) {
                    const int k = kc * kBK + c * 4;
                    const float m0 = s_mean[k], m1 = s_mean[k + 1], m2 = s_mean[k + 2], m3 = s_mean[k + 3];
                    const float r0 = s_rstd[k], r1 = s_rstd[k + 1], r2 = s_rstd[k + 2], r3 = s_rstd[k + 3];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (valid[i]) {
                            cur[i].x = (cur[i].x - m0) * r0;
                            cur[i].y = (cur[i].y - m1) * r1;
                            cur[i].z = (cur[i].z - m2) * r2;
                            cur[i].w = (cur[i].w - m3) * r3;
                        }
                    }
                }
                if (kc + 2 * kAGroups < nk) load_chunk(buf, kc + 2 * kAGroups);
                const int g = gbase + kc;
                const int s = g % kStages;
                const uint32_t ph = (g / kStages) & 1;
                if (lane == 0) mbar_wait(empty_bar(s), ph ^ 1);
                __syncwarp();
                uint8_t *stage = smem + (size_t)s * S::kStage;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 hi, lo;
                    split_tf32(cur[i].x, hi.x, lo.x);
                    split_tf32(cur[i].y, hi.y, lo.y);
                    split_tf32(cur[i].z, hi.z, lo.z);
                    split_tf32(cur[i].w, hi.w, lo.w);
                    const uint32_t off = sw64_offset(rsub + 32 * i, c);
                    *reinterpret_cast<float4 *>(stage + off) = hi;
                    *reinterpret_cast<float4 *>(stage + S::kAHalf + off) = lo;
                }
                fence_async_smem();
                mbar_arrive(full_bar(s));
            };
            float4 buf0[4], buf1[4];   // this group's chunks grp, grp + 2, grp + 4, ...
            if (grp < nk) load_chunk(buf0, grp);
            if (grp + kAGroups < nk) load_chunk(buf1, grp + kAGroups);
            for (int kc = grp; kc < nk; kc += 2 * kAGroups) {
                produce(buf0, kc);
                if (kc + kAGroups < nk) produce(buf1, kc + kAGroups);
            }
            gbase += nk;
        }
    } else if (warp == 8) {
        // ================= B producer: this rank's half of every weight stage =======================================
        if (lane == 0) {
            int g = 0;
            constexpr int kHalfElems = (BN / 2) * kBK;
            for (int t = cluster_id; t < total; t += n_clusters) {
                int f, ct;
                long long row0, n_rows;
                if (!decode(t, f, ct, row0, n_rows)) continue;
                const float *src = wpack + (((size_t)ct * nk) * 2 + rank) * (2 * kHalfElems);
                for (int kc = 0; kc < nk; ++kc, ++g) {
                    const int s = g % kStages;
                    const uint32_t ph = (g / kStages) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    mbar_arrive_expect_tx(full_bar(s), 2 * S::kBHalf);
                    bulk_g2s(sbase + s * S::kStage + 2 * S::kAHalf, src + (size_t)kc * 2 * (2 * kHalfElems), 2 * S::kBHalf, full_bar(s));
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0 && rank == 0) {
            // ================= MMA issuer (leader CTA) ================================================================
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((256u >> 4) << 24);
            int g = 0, it = 0;
            for (int t = cluster_id; t < total; t += n_clusters) {
                int f, ct;
                long long row0, n_rows;
                if (!decode(t, f, ct, row0, n_rows)) continue;
                const int ab = it & 1;
                if (!mma_only) mbar_wait(tmem_empty_bar(ab), ((it >> 1) & 1) ^ 1);  // both CTAs drained this accumulator buffer
                tc_fence_after();
                for (int kc = 0; kc < nk; ++kc, ++g) {
                    const int s = g % kStages;
                    const uint32_t ph = (g / kStages) & 1;
                    if (!mma_only) {
                        mbar_wait(full_bar(s), ph);
                        mbar_wait(peer_full_bar(s), ph);
                    }   // plain (cta-scope) wait as CUTLASS's 2-SM pipelines do: the data is read by
                                                       // the async proxy, ordered by the writers' fence.proxy.async + release.cluster arrive
                    tc_fence_after();
                    const uint32_t sA = sbase + s * S::kStage, sB = sA + 2 * S::kAHalf;
                    const uint32_t d = tmem_base + ab * BN;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint64_t a_hi = make_desc(sA + ks * 32), a_lo = make_desc(sA + S::kAHalf + ks * 32);
                        const uint64_t b_hi = make_desc(sB + ks * 32), b_lo = make_desc(sB + S::kBHalf + ks * 32);
                        mma2_tf32(d, a_lo, b_hi, idesc, (kc | ks) != 0);
                        mma2_tf32(d, a_hi, b_lo, idesc, 1);
                        mma2_tf32(d, a_hi, b_hi, idesc, 1);
                    }
                    mma2_commit_multicast(empty_bar(s), 3);  // frees stage s in BOTH CTAs
                }
                mma2_commit_multicast(accum_bar(ab), 3);
                ++it;
            }
        } else if (rank != 0 && lane < kStages && !mma_only) {
            // ================= relay (peer CTA): lane L forwards "stage L is full" to the leader's peer_full[L] ==========
            // one lane per stage, so the wake-up + remote-arrive latency of consecutive stages overlaps
            int gbase = 0;
            for (int t = cluster_id; t < total; t += n_clusters) {
                int f, ct;
                long long row0, n_rows;
                if (!decode(t, f, ct, row0, n_rows)) continue;
                for (int kc = 0; kc < nk; ++kc) {
                    const int g = gbase + kc;
                    if (g % kStages != lane) continue;
                    mbar_wait(full_bar(lane), (g / kStages) & 1);
                    mbar_arrive_remote_relaxed(map_to_cta(peer_full_bar(lane), 0));
                }
                gbase += nk;
            }
        }
    } else {
        // ================= epilogue warps 10..13: TMEM lane quarter q = warp % 4 =====================================
        const int q = warp & 3, ew = warp - 10, et = tid - 10 * 32;
        int it = 0;
        for (int t = cluster_id; t < total; t += n_clusters) {
            int f, ct;
            long long row0, n_rows;
            if (!decode(t, f, ct, row0, n_rows)) continue;
            row0 += (long long)rank * kRowsPerCta;
            const int n0 = ct * BN, ab = it & 1;
            double acc_s[BN / 32], acc_ss[BN / 32];
#pragma unroll
            for (int cb = 0; cb < BN / 32; ++cb) acc_s[cb] = 0.0, acc_ss[cb] = 0.0;
            named_bar_sync(2, 128);
            for (int i = et; i < 4 * BN * 2; i += 128) s_part[i] = 0.0;
            named_bar_sync(2, 128);
            mbar_wait(accum_bar(ab), (it >> 1) & 1);
            tc_fence_after();
            const long long r = row0 + q * 32 + lane;
            const bool valid = r < n_rows;
            const float w = valid ? (a.row_w ? a.row_w[(size_t)f * a.rowcap + r] : 1.f) : 0.f;
            const float m = w == 1.f ? 1.f : 0.f;
            const bool heavy = w != 0.f && w != 1.f;
            float *yrow = a.Y ? a.Y + ((size_t)f * a.rowcap + r) * a.ldy + n0 : nullptr;
#pragma unroll 1
            for (int cb = 0; cb < ((a.dbg & 2) ? 0 : BN / 32); ++cb) {
                float v[32], p2[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ab * BN + cb * 32, v);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.bias + n0 + cb * 32 + j));
                    v[j] = fmaxf(v[j] + b4.x, 0.f);
                    v[j + 1] = fmaxf(v[j + 1] + b4.y, 0.f);
                    v[j + 2] = fmaxf(v[j + 2] + b4.z, 0.f);
                    v[j + 3] = fmaxf(v[j + 3] + b4.w, 0.f);
                }
                if (yrow && valid && !(a.dbg & 16)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(yrow + cb * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                if (a.dbg & 32) continue;
                if (heavy) {
                    for (int j = 0; j < 32; ++j) {
                        const double y = (double)v[j], wy = (double)w * y;
                        atomicAdd(&s_part[((size_t)ew * BN + cb * 32 + j) * 2], wy);
                        atomicAdd(&s_part[((size_t)ew * BN + cb * 32 + j) * 2 + 1], wy * y);
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] *= m;
                    p2[j] = v[j] * v[j];
                }
                acc_s[cb] += (double)butterfly_colsum(v, lane);
                acc_ss[cb] += (double)butterfly_colsum(p2, lane);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {  // this warp has drained its quarter of accumulator buffer ab
                if (rank == 0) mbar_arrive(tmem_empty_bar(ab));
                else mbar_arrive_remote(map_to_cta(tmem_empty_bar(ab), 0));
            }
#pragma unroll
            for (int cb = 0; cb < BN / 32; ++cb) {
                atomicAdd(&s_part[((size_t)ew * BN + cb * 32 + lane) * 2], acc_s[cb]);
                atomicAdd(&s_part[((size_t)ew * BN + cb * 32 + lane) * 2 + 1], acc_ss[cb]);
            }
            named_bar_sync(2, 128);
            for (int i = et; i < BN * 2; i += 128) {
                const double s4 = s_part[i] + s_part[BN * 2 + i] + s_part[2 * BN * 2 + i] + s_part[3 * BN * 2 + i];
                atomicAdd(a.out_stats + ((size_t)f * a.Cout + n0) * 2 + i, s4);
            }
            ++it;
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody leaves while the peer can still touch its barriers, shared memory or TMEM
    if (warp == 9) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

template <int BN>
int launch_tc2(const LayerArgs &a, int F, float *wpack, cudaStream_t st) {
    using S = Smem2<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        MVX_CUDA_CHECK(cudaFuncSetAttribute(tc2_layer_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        attr_set = true;
    }
    const int total = a.Cin * a.Cout;
    pack_weights2_kernel<BN><<<(total + 255) / 256, 256, 0, st>>>(a.Wt, a.Cin, a.Cout, wpack);
    MVX_LAUNCH_CHECK();
    const long long max_rows = a.counts ? a.rowcap : a.rows_fixed;
    const int row_tiles = (int)ceil_div(max_rows, 2 * kRowsPerCta), col_tiles = a.Cout / BN;
    const long long slots = (long long)F * row_tiles * col_tiles;
    const int clusters = (int)(slots < kSMs / 2 ? slots : kSMs / 2);
    tc2_layer_kernel<BN><<<2 * clusters, kThreads2, S::kTotal, st>>>(a, wpack, F, row_tiles, col_tiles);
    MVX_LAUNCH_CHECK();
    return MVX_OK;
}

}  // namespace

bool tc2_layer_eligible(const LayerArgs &a) {
    return a.vmax == nullptr && a.Cin % kBK == 0 && a.Cin <= 768 && a.Cout % 128 == 0 && a.ldx % 4 == 0 &&
           (a.Y == nullptr || a.ldy % 4 == 0);
}

int launch_layer_tc2(const LayerArgs &a_in, int F, float *wpack, cudaStream_t st) {
    LayerArgs a = a_in;
    if (const char *e = getenv("MVX_DBG")) a.dbg = atoi(e);
    MVX_REQUIRE(tc2_layer_eligible(a) && wpack, MVX_EINVAL, "layer not eligible for the CTA-pair tensor-core kernel");
    const long long max_rows = a.counts ? a.rowcap : a.rows_fixed;
    if (max_rows <= 0) return MVX_OK;
    if (a.Cout % 256 == 0) return launch_tc2<256>(a, F, wpack, st);
    return launch_tc2<128>(a, F, wpack, st);
}

}  // namespace mvx
