// NM_PREC_FP32: the same fused forward (src/nerf/models.py:60-80) on the CUDA cores, plain fp32 FMA arithmetic.
// It exists as the GPU-side numerical yard-stick for the tensor-core kernel (its only difference from the reference
// is fp32 summation order) and is selectable through NmRenderCfg.precision; it is not the performance path.
//
// One CTA = 64 points, 256 threads; activations ping-pong between two [256][64] shared-memory panels (feature-major
// so that the 8 points a thread owns are bank-conflict free); each thread accumulates an 8-point x 8-output
// register tile; weights come transposed (Wt[k][n], nm_program.cu) through the read-only path.
#include "nm_common.h"
#include "nm_frontend.cuh"

namespace nm {
namespace {

constexpr int kPts = 64;
constexpr int kThreads = 256;

struct SimtParams {
  const NetProgram* prog;
  const float* wt;
  const float* bias;
  const float* head;
  MlpInput in;
  float* out;
  int out_sigma_only;
};

__global__ void __launch_bounds__(kThreads, 1) mlp_simt_kernel(const __grid_constant__ SimtParams P) {
  extern __shared__ float sm[];
  float* X0 = sm;                    // [256][64]
  float* X1 = X0 + 256 * kPts;       // [256][64]
  float* PEX = X1 + 256 * kPts;      // [64][64]
  float* PED = PEX + 64 * kPts;      // [64][64]
  float* SIG = PED + 64 * kPts;      // [64]
  const NetProgram& G = *P.prog;
  const int tid = threadIdx.x;
  const int tn = tid >> 3, tp = tid & 7;
  const long long m0 = (long long)blockIdx.x * kPts;

  if (tid < kPts) {
    long long m = m0 + tid;
    if (m >= P.in.M) m = P.in.M - 1;
    float p[3], d[3];
    fetch_point(P.in, m, p, d);
    positional_encoding(p, G.L_xyz, G.inc_xyz, G.freq_xyz, [&](int j, float v) { PEX[j * kPts + tid] = v; });
    if (G.dim_dir > 0)
      positional_encoding(d, G.L_dir, G.inc_dir, G.freq_dir, [&](int j, float v) { PED[j * kPts + tid] = v; });
  }
  __syncthreads();

  float* cur = X0;
  float* nxt = X1;
  for (int li = 0; li < G.n_layers; ++li) {
    const LayerProg L = G.layers[li];
    const int N = L.n_out;
    const int n0 = tn * 8;
    const float* Wt = P.wt + L.wt_off;
    if (n0 < N) {
      float acc[8][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float b = __ldg(P.bias + L.bias_off + n0 + j);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][j] = b;
      }
      auto segment = [&](const float* X, int K, const float* W) {
        for (int k = 0; k < K; ++k) {
          float a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = X[k * kPts + tp + 8 * i];
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + (size_t)k * N + n0));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + (size_t)k * N + n0 + 4));
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
      };
      segment(cur, L.k_act, Wt);                                   // input order [activations | encoding]
      if (L.pe_src) segment(L.pe_src == SRC_PE_XYZ ? PEX : PED, L.k_pe, Wt + (size_t)L.k_act * N);
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v = acc[i][j];
          if (L.relu) v = fmaxf(v, 0.f);
          nxt[(n0 + j) * kPts + tp + 8 * i] = v;
        }
    }
    __syncthreads();
    // heads: one thread per point
    if (L.kind != KIND_HIDDEN && tid < kPts) {
      const int heads = L.kind == KIND_SIGMA ? 1 : (L.kind == KIND_RGB ? 3 : 4);
      const float* hw = P.head + L.head_off;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      for (int hh = 0; hh < heads; ++hh) {
        float s = 0.f;
        for (int k = 0; k < N; ++k) s = fmaf(__ldg(hw + hh * N + k), nxt[k * kPts + tid], s);
        o[hh] = s + __ldg(hw + heads * N + hh);
      }
      const long long m = m0 + tid;
      if (L.kind == KIND_SIGMA) {
        SIG[tid] = o[0];
        if (L.is_final && m < P.in.M) P.out[m] = o[0];
      } else if (m < P.in.M) {
        const float sg = (L.kind == KIND_RGB) ? SIG[tid] : o[3];
        if (P.out_sigma_only) {
          P.out[m] = sg;
        } else {
          float4 r;
          r.x = 1.f / (1.f + expf(-o[0]));
          r.y = 1.f / (1.f + expf(-o[1]));
          r.z = 1.f / (1.f + expf(-o[2]));
          r.w = sg;
          reinterpret_cast<float4*>(P.out)[m] = r;
        }
      }
    }
    float* t = cur; cur = nxt; nxt = t;
  }
}

}  // namespace

int launch_mlp_simt(const NetDev& net, bool sigma_only, const MlpInput& in, float* out, cudaStream_t st,
                    int64_t* launches) {
  if (in.M <= 0) return 0;
  SimtParams P{};
  P.prog = sigma_only ? net.d_sigma : net.d_full;
  P.wt = net.d_wt;
  P.bias = net.d_bias;
  P.head = net.d_head;
  P.in = in;
  P.out = out;
  P.out_sigma_only = sigma_only ? 1 : 0;
  const size_t smem = (size_t)(2 * 256 * kPts + 2 * 64 * kPts + kPts) * sizeof(float);
  NM_CUDA(cudaFuncSetAttribute(mlp_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per current device
  const long long grid = (in.M + kPts - 1) / kPts;
  NM_CHECK(grid < (1ll << 31), "too many points for one launch");
  mlp_simt_kernel<<<(unsigned)grid, kThreads, smem, st>>>(P);
  NM_CUDA(cudaGetLastError());
  if (launches) ++*launches;
  return 0;
}

}  // namespace nm
