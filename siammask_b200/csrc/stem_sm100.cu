// Stem of the backbone on the tensor cores: conv 7x7 stride 2 pad 0, 3 -> 64 channels, + BN + ReLU
// (experiments/siammask_sharp/resnet.py:154,218-220), reading the raw NCHW fp32 pixels the tracker loop hands
// over (tools/test.py:61-64) and writing p0 as NHWC split-fp16 planes.
//
// Cin = 3 is useless to TMA (6-byte pixels), so the A operand is BUILT: four producer warps gather each output
// pixel's 7x7x3 window straight from global memory (L1/L2 resident: neighbouring windows overlap 5/7), split it
// into fp16 hi/lo and write it into the 128B-swizzled K-major tile layout tcgen05.mma expects
// (K = 147 padded to 192 = three 64-wide k-blocks, a 4-deep ring).  The 64 x 192 weight matrix stays resident in
// shared memory for the whole persistent CTA.  MMA issue, TMEM double buffering and the smem + TMA-store
// epilogue are the ones of conv_gemm_sm100.cu with N = 64.
//
// Warps (448 threads): 0 = weight loader, 1 = TMEM owner + MMA issuer, 2..5 = A producers (one tile row per
// thread), 6..13 = epilogue (two per TMEM lane quarter).
#include "common.cuh"
#include "ptx.cuh"

namespace smk {

namespace {

constexpr int BM = 128, BN = 64, BK = 64, UK = 16;
constexpr int KBLOCKS = 3;                 // K = 7*7*3 = 147 -> 192
constexpr int KREAL = 147;
constexpr int STAGES = 4;
constexpr int A_TILE = BM * BK * 2;        // 16 KB
constexpr int B_TILE = BN * BK * 2;        // 8 KB
constexpr int STG_TILE = 32 * 64;          // 32 rows x 32 cols fp16
constexpr int NPROD = 256;                 // two producer threads per tile row (each builds half of every k-block:
                                           // the gather is a per-thread latency chain of 147 loads + conversions)
constexpr int NEPI = 8;
constexpr int NTHREADS = 64 + NPROD + 32 * NEPI;

struct StemParams {
  CUtensorMap tmB[2];     // weights [64][192] fp16 K-major, box 64 x 64, hi / lo
  CUtensorMap tmOut[2];   // p0 planes [M][64], box 32 x 32, 64B swizzle
  const float* x;         // [B][3][S][S]
  const float* alpha;     // [64] 2^-e
  const float* beta;      // [64]
  int* ovf;               // overflow flag (values outside fp16's range), may be null
  int B, S, So, M, m_tiles;
};

template <int NSPLIT>
struct SCfg {
  static constexpr int STAGE_BYTES = NSPLIT * A_TILE;
  static constexpr int B_BYTES = KBLOCKS * NSPLIT * B_TILE;
  static constexpr int STG_BYTES = NEPI * NSPLIT * STG_TILE;
  static constexpr int SMEM = B_BYTES + STAGES * STAGE_BYTES + STG_BYTES + 1024 + 256;
  static constexpr int ACC_COLS = NSPLIT * BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 128 ? 128 : 256;
  static_assert(SMEM <= 227 * 1024, "smem budget");
};

// one 64-wide k-block of this thread's row: k = (r*7 + s)*3 + c
// JH = which half of the k-block (4 of its 8 16-byte chunks) this thread builds; compile-time so that k -> (r, s, c)
// folds into constant offsets
template <int KB, int NSPLIT, int JH>
__device__ __forceinline__ void build_kblock(const float* __restrict__ base, bool valid, int S, uint8_t* stage,
                                             int row) {
  uint8_t* rowp = stage + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    constexpr int J0 = JH * 4;
    const int j = J0 + jj;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = KB * 64 + j * 8 + e;
      if (k < KREAL) {
        const int c = k % 3, rs = k / 3, r = rs / 7, s = rs % 7;
        v[e] = valid ? __ldg(base + ((size_t)c * S + r) * S + s) : 0.f;
      } else {
        v[e] = 0.f;
      }
    }
    uint4 h, l;
    __half2* hh = reinterpret_cast<__half2*>(&h);
    __half2* ll = reinterpret_cast<__half2*>(&l);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const __half2 hv = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
      hh[t] = hv;
      const float2 hf = __half22float2(hv);
      ll[t] = __floats2half2_rn(v[2 * t] - hf.x, v[2 * t + 1] - hf.y);
    }
    const int off = (j ^ (row & 7)) << 4;       // SWIZZLE_128B: 16B chunk index ^= row & 7
    *reinterpret_cast<uint4*>(rowp + off) = h;
    if constexpr (NSPLIT == 2) *reinterpret_cast<uint4*>(rowp + A_TILE + off) = l;
  }
}

template <int NSPLIT>
__global__ void __launch_bounds__(NTHREADS, 1) stem_tc_kernel(const __grid_constant__ StemParams p) {
  using C = SCfg<NSPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_smem = smem;                                  // [kb][plane][64 x 128B]
  uint8_t* a_smem = b_smem + C::B_BYTES;                   // ring of [plane][128 x 128B]
  uint8_t* stg_base = a_smem + STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + C::STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* b_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_bar + 1);

  // shuffle-broadcast warp index + elected issuing lane: see conv_gemm_sm100.cu (keeps tcgen05 / TMA issue convergent)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const bool leader = elect_one();

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSPLIT; ++i) { tma_prefetch_desc(&p.tmB[i]); tma_prefetch_desc(&p.tmOut[i]); }
    // one arrival per producer WARP (after its lanes' fences): 256 single-thread arrivals serialise on the barrier word
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], NPROD / 32); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], NEPI); }
    mbar_init(b_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) { __syncwarp(); tmem_alloc<C::TMEM_COLS>(tmem_slot); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  // broadcast through a shuffle so the compiler KNOWS the TMEM base is warp-uniform: tcgen05 operands live in uniform
  // registers, and a value that merely came out of shared memory makes ptxas wrap every single MMA in a
  // divergence ("waterfall") loop — ELECT / R2UR / branch per instruction, ~90 clk of issue time per MMA
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // resident weights: 3 k-blocks x NSPLIT planes
    if (leader) {
      mbar_arrive_expect_tx(b_bar, C::B_BYTES);
      for (int kb = 0; kb < KBLOCKS; ++kb)
        for (int s = 0; s < NSPLIT; ++s)
          tma_load_2d(b_smem + (kb * NSPLIT + s) * B_TILE, &p.tmB[s], b_bar, kb * BK, 0);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
    mbar_wait(b_bar, 0);
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * C::ACC_COLS;
#pragma unroll 1
      for (int kb = 0; kb < KBLOCKS; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t a_hi = smem_u32(a_smem + stage * C::STAGE_BYTES);
        const uint32_t b_hi = smem_u32(b_smem + kb * NSPLIT * B_TILE);
        const uint64_t da_hi0 = umma_desc_kmajor_sw128(a_hi), db_hi0 = umma_desc_kmajor_sw128(b_hi);
        const uint64_t da_lo0 = umma_desc_kmajor_sw128(a_hi + A_TILE), db_lo0 = umma_desc_kmajor_sw128(b_hi + B_TILE);
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t kadd = static_cast<uint64_t>(k * UK * 2 / 16);     // 32 B per K step in the address field
            umma_f16(tmem_d, da_hi0 + kadd, db_hi0 + kadd, idesc, (kb | k) != 0 ? 1u : 0u);
            if constexpr (NSPLIT == 2) {
              umma_f16(tmem_d + BN, da_lo0 + kadd, db_hi0 + kadd, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_f16(tmem_d + BN, da_hi0 + kadd, db_lo0 + kadd, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (kb == KBLOCKS - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2 && warp < 2 + NPROD / 32) {
    // ===================== A producers: one output pixel (tile row) per thread =====================
    const int row = (threadIdx.x - 64) & (BM - 1);
    const bool upper = threadIdx.x - 64 >= BM;          // warp-uniform: warps 2-5 build chunks 0-3, warps 6-9 chunks 4-7
    int stage = 0;
    uint32_t phase = 0;
    const int So2 = p.So * p.So;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      const int m = tile * BM + row;
      const bool valid = m < p.M;
      const int b = m / So2;
      const int rem = m - b * So2;
      const int y = rem / p.So;
      const int xo = rem - y * p.So;
      const float* base = p.x + ((size_t)b * 3 * p.S + 2 * y) * p.S + 2 * xo;
#define SMK_STEM_KB(KB)                                                          \
  mbar_wait(&empty_bar[stage], phase ^ 1);                                       \
  if (upper) build_kblock<KB, NSPLIT, 1>(base, valid, p.S, a_smem + stage * C::STAGE_BYTES, row); \
  else build_kblock<KB, NSPLIT, 0>(base, valid, p.S, a_smem + stage * C::STAGE_BYTES, row); \
  fence_proxy_async();                                                           \
  __syncwarp();                                                                  \
  if (lane == 0) mbar_arrive(&full_bar[stage]);                                  \
  if (++stage == STAGES) { stage = 0; phase ^= 1; }
      SMK_STEM_KB(0)
      SMK_STEM_KB(1)
      SMK_STEM_KB(2)
#undef SMK_STEM_KB
    }
  } else if (warp >= 2 + NPROD / 32) {
    // ===================== epilogue: +beta, ReLU, split, smem, TMA store =====================
    const int ew = warp - (2 + NPROD / 32);       // 0..7
    const int quarter = warp & 3;                 // TMEM lane quarter = warp id % 4
    // the quarter's two warps: take 32-column chunk 0 / 1
    const int chunk = (ew >> 2) & 1;
    uint8_t* buf = stg_base + ew * NSPLIT * STG_TILE;
    const int swz = (lane >> 1) & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * C::ACC_COLS + chunk * 32;
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr, r);
      if constexpr (NSPLIT == 2) {
        uint32_t r2[32];
        tmem_ld_32x32b_x32(taddr + BN, r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
      } else {
        tmem_ld_wait();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);      // accumulator is in registers: release it early
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 al = __ldg(reinterpret_cast<const float4*>(p.alpha + chunk * 32 + j));
        const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + chunk * 32 + j));
        v[j + 0] = fmaxf(fmaf(__uint_as_float(r[j + 0]), al.x, be.x), 0.f);
        v[j + 1] = fmaxf(fmaf(__uint_as_float(r[j + 1]), al.y, be.y), 0.f);
        v[j + 2] = fmaxf(fmaf(__uint_as_float(r[j + 2]), al.z, be.z), 0.f);
        v[j + 3] = fmaxf(fmaf(__uint_as_float(r[j + 3]), al.w, be.w), 0.f);
      }
      {
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) amax = fmaxf(amax, v[j]);
        flag_if_out_of_range(amax, p.ovf);
      }
      if (leader) tma_store_wait_read<0>();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 h, l;
        __half2* hh = reinterpret_cast<__half2*>(&h);
        __half2* ll = reinterpret_cast<__half2*>(&l);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const __half2 hv = __floats2half2_rn(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]);
          hh[t] = hv;
          const float2 hf = __half22float2(hv);
          ll[t] = __floats2half2_rn(v[8 * j + 2 * t] - hf.x, v[8 * j + 2 * t + 1] - hf.y);
        }
        *reinterpret_cast<uint4*>(buf + lane * 64 + ((j ^ swz) << 4)) = h;
        if constexpr (NSPLIT == 2) *reinterpret_cast<uint4*>(buf + STG_TILE + lane * 64 + ((j ^ swz) << 4)) = l;
      }
      fence_proxy_async();
      __syncwarp();
      if (leader) {
#pragma unroll
        for (int s = 0; s < NSPLIT; ++s) tma_store_2d(&p.tmOut[s], buf + s * STG_TILE, chunk * 32, tile * BM + quarter * 32);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

}  // namespace

CUtensorMap make_map_2d_any(const __half* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer,
                            int swizzle_bytes);

void launch_stem_tc(const float* x, int B, int S, const __half* w_hi, const __half* w_lo, const float* alpha,
                    const float* beta, Act out, int num_sms, cudaStream_t st, int* ovf) {
  const int So = (S - 7) / 2 + 1;
  SMK_CHECK(out.H == So && out.W == So && out.C == 64 && out.B == B, "stem output shape");
  StemParams p;
  p.ovf = ovf;
  p.x = x;
  p.alpha = alpha;
  p.beta = beta;
  p.B = B;
  p.S = S;
  p.So = So;
  p.M = B * So * So;
  p.m_tiles = (p.M + BM - 1) / BM;
  const int nsplit = out.lo != nullptr ? 2 : 1;
  for (int s = 0; s < nsplit; ++s) {
    p.tmB[s] = make_map_2d_any(s == 0 ? w_hi : w_lo, KBLOCKS * BK, BN, BK, BN, 128);
    p.tmOut[s] = make_map_2d_any(s == 0 ? out.hi : out.lo, 64, (uint64_t)p.M, 32, 32, 64);
  }
  if (nsplit == 1) { p.tmB[1] = p.tmB[0]; p.tmOut[1] = p.tmOut[0]; }
  const int grid = p.m_tiles < num_sms ? p.m_tiles : num_sms;
  if (nsplit == 2) {
    static unsigned long long attr = 0;
    ensure_dynamic_smem(stem_tc_kernel<2>, SCfg<2>::SMEM, attr);
    stem_tc_kernel<2><<<grid, NTHREADS, SCfg<2>::SMEM, st>>>(p);
  } else {
    static unsigned long long attr = 0;
    ensure_dynamic_smem(stem_tc_kernel<1>, SCfg<1>::SMEM, attr);
    stem_tc_kernel<1><<<grid, NTHREADS, SCfg<1>::SMEM, st>>>(p);
  }
  SMK_CUDA(cudaGetLastError());
}

}  // namespace smk
