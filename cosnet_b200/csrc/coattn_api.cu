// Host side of the C ABI declared in include/coattn_b200.h: argument checks, workspace carving,
// TMA descriptor encoding and kernel launches.  No torch types, no global mutable state.
#include "../../include/coattn_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

#include "coattn_kernels.cuh"
#include "attend2_kernel.cuh"

namespace {

using namespace coattn;

constexpr int64_t kAlign = 1024;
inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Layout {
  int N, L, Lp;
  int64_t off_t, off_at, off_vv, off_w16, off_z, off_lse, total;
  int64_t bytes_t, bytes_at, bytes_vv, bytes_w16, bytes_z, bytes_lse;
  // element strides
  int64_t t_pass_elems() const { return (int64_t)N * Lp * kC; }   // T[pass] and VV[pass]
};

Layout make_layout(int n, int h, int w) {
  Layout ly{};
  ly.N = n;
  ly.L = h * w;
  ly.Lp = (int)round_up(ly.L, kLPad);
  const int64_t plane = (int64_t)n * ly.Lp * kC * 2;  // one bf16 [N][Lp][C] (or [N][C][Lp]) array
  int64_t off = 0;
  ly.off_t = off;   ly.bytes_t = 2 * plane;   off = round_up(off + ly.bytes_t, kAlign);
  ly.off_at = off;  ly.bytes_at = plane;      off = round_up(off + ly.bytes_at, kAlign);
  ly.off_vv = off;  ly.bytes_vv = 2 * plane;  off = round_up(off + ly.bytes_vv, kAlign);
  ly.off_w16 = off; ly.bytes_w16 = (int64_t)kC * kC * 2; off = round_up(off + ly.bytes_w16, kAlign);
  ly.off_z = off;   ly.bytes_z = (int64_t)2 * n * kC * ly.L * 4; off = round_up(off + ly.bytes_z, kAlign);
  ly.off_lse = off; ly.bytes_lse = (int64_t)2 * n * ly.L * 4;    off = round_up(off + ly.bytes_lse, kAlign);
  ly.total = off;
  return ly;
}

int check_dims(int n, int c, int h, int w) {
  if (n < 1 || h < 1 || w < 1 || c != kC) return COATTN_E_SHAPE;
  if ((int64_t)h * w > (1 << 20)) return COATTN_E_SHAPE;
  return COATTN_OK;
}

int check_workspace(const void* ws, int64_t bytes, const Layout& ly) {
  if (!ws) return COATTN_E_NULL;
  if ((reinterpret_cast<uintptr_t>(ws) & (kAlign - 1)) != 0) return COATTN_E_WORKSPACE;
  if (bytes < ly.total) return COATTN_E_WORKSPACE;
  return COATTN_OK;
}

int check_arch(int* sm_count) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return (int)e;
  if (major != 10) return COATTN_E_ARCH;
  if (sm_count) {
    e = cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
  }
  return COATTN_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  // resolved through the runtime so the library has no link-time dependency on libcuda
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess) return nullptr;
  if (qres != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<EncodeTiledFn>(fn);
}

// 2-D 16-bit row-major tensor [rows][cols]; box = {64 columns (128 B), box_rows}; 128-byte swizzle.
int make_tmap(EncodeTiledFn enc, CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
              uint32_t box_rows, bool bf16) {
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? COATTN_OK : COATTN_E_DRIVER;
}

inline uint8_t* seg(void* ws, int64_t off) { return static_cast<uint8_t*>(ws) + off; }

}  // namespace

extern "C" {

int coattn_b200_abi_version(void) { return COATTN_B200_ABI_VERSION; }

const char* coattn_b200_strerror(int code) {
  switch (code) {
    case COATTN_OK: return "ok";
    case COATTN_E_NULL: return "required pointer is NULL";
    case COATTN_E_SHAPE: return "bad shape (need n,h,w >= 1 and c == 256)";
    case COATTN_E_WORKSPACE: return "workspace too small or not 1024-byte aligned";
    case COATTN_E_ARCH: return "device is not sm_100 class (B200); there is no fallback path";
    case COATTN_E_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    case COATTN_E_ALIGN: return "tensor pointer not 16-byte aligned";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

int64_t coattn_workspace_bytes(int n, int c, int h, int w) {
  if (check_dims(n, c, h, w) != COATTN_OK) return COATTN_E_SHAPE;
  return make_layout(n, h, w).total;
}

int coattn_workspace_segment(const char* name, int n, int c, int h, int w, int64_t* offset, int64_t* bytes) {
  if (!name || !offset || !bytes) return COATTN_E_NULL;
  if (int e = check_dims(n, c, h, w)) return e;
  const Layout ly = make_layout(n, h, w);
  const int64_t plane = (int64_t)n * ly.Lp * kC * 2;
  if (!strcmp(name, "bt")) { *offset = ly.off_t; *bytes = plane; }
  else if (!strcmp(name, "qt")) { *offset = ly.off_t + plane; *bytes = plane; }
  else if (!strcmp(name, "at")) { *offset = ly.off_at; *bytes = plane; }
  else if (!strcmp(name, "b16")) { *offset = ly.off_vv; *bytes = plane; }
  else if (!strcmp(name, "a16")) { *offset = ly.off_vv + plane; *bytes = plane; }
  else if (!strcmp(name, "w16")) { *offset = ly.off_w16; *bytes = ly.bytes_w16; }
  else if (!strcmp(name, "z")) { *offset = ly.off_z; *bytes = ly.bytes_z; }
  else if (!strcmp(name, "lse")) { *offset = ly.off_lse; *bytes = ly.bytes_lse; }
  else return COATTN_E_NULL;
  return COATTN_OK;
}

int coattn_stage_prep(const float* v_a, const float* v_b, const float* w, void* workspace,
                      int64_t workspace_bytes, int n, int c, int h, int w_, unsigned flags, void* stream) {
  if (!v_a || !v_b || !w) return COATTN_E_NULL;
  if (int e = check_dims(n, c, h, w_)) return e;
  const Layout ly = make_layout(n, h, w_);
  if (int e = check_workspace(workspace, workspace_bytes, ly)) return e;
  if (int e = check_arch(nullptr)) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t plane_elems = ly.t_pass_elems();
  PrepParams p;
  p.va = v_a;
  p.vb = v_b;
  p.bt = reinterpret_cast<unsigned short*>(seg(workspace, ly.off_t));
  p.at = reinterpret_cast<unsigned short*>(seg(workspace, ly.off_at));
  p.b16 = reinterpret_cast<unsigned short*>(seg(workspace, ly.off_vv));
  p.a16 = p.b16 + plane_elems;
  p.L = ly.L;
  p.Lp = ly.Lp;
  unsigned short* w16 = reinterpret_cast<unsigned short*>(seg(workspace, ly.off_w16));
  const dim3 grid(ly.Lp / kPrepTileL, 2 * n);
  const bool vec = (ly.L % 4 == 0) &&
                   (((reinterpret_cast<uintptr_t>(v_a) | reinterpret_cast<uintptr_t>(v_b)) & 15) == 0);
  if (flags & COATTN_FLAG_BF16) {
    if (vec) prep_kernel_vec4<true><<<grid, kPrepThreads, 0, st>>>(p);
    else prep_kernel<true><<<grid, kPrepThreads, 0, st>>>(p);
    cast_w_kernel<true><<<(kC * kC + 255) / 256, 256, 0, st>>>(w, w16, kC * kC);
  } else {
    if (vec) prep_kernel_vec4<false><<<grid, kPrepThreads, 0, st>>>(p);
    else prep_kernel<false><<<grid, kPrepThreads, 0, st>>>(p);
    cast_w_kernel<false><<<(kC * kC + 255) / 256, 256, 0, st>>>(w, w16, kC * kC);
  }
  return (int)cudaGetLastError();
}

int coattn_stage_project(void* workspace, int64_t workspace_bytes, int n, int c, int h, int w_, unsigned flags,
                         void* stream) {
  const bool bf16 = (flags & COATTN_FLAG_BF16) != 0;
  if (int e = check_dims(n, c, h, w_)) return e;
  const Layout ly = make_layout(n, h, w_);
  if (int e = check_workspace(workspace, workspace_bytes, ly)) return e;
  if (int e = check_arch(nullptr)) return e;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return COATTN_E_DRIVER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tm_at, tm_w;
  if (int e = make_tmap(enc, &tm_at, seg(workspace, ly.off_at), (uint64_t)n * ly.Lp, kC, 128, bf16)) return e;
  if (int e = make_tmap(enc, &tm_w, seg(workspace, ly.off_w16), kC, kC, 256, bf16)) return e;
  ProjectParams p;
  p.qt = reinterpret_cast<unsigned short*>(seg(workspace, ly.off_t)) + ly.t_pass_elems();
  p.Lp = ly.Lp;
  auto kern = bf16 ? project_kernel<true> : project_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kProjSmemBytes);
  if (e != cudaSuccess) return (int)e;
  kern<<<dim3(ly.Lp / 128, n), kNumThreads, kProjSmemBytes, st>>>(tm_at, tm_w, p);
  return (int)cudaGetLastError();
}

static int launch_attend(const float* v_a, const float* v_b, float* cat_a, float* cat_b, float* z, float* lse,
                         float* mask, const float* gate_w, const float* gate_b, void* workspace, int64_t workspace_bytes, int n, int c, int h, int w_,
                         unsigned flags, void* stream) {
  const bool bf16 = (flags & COATTN_FLAG_BF16) != 0;
  if (int e = check_dims(n, c, h, w_)) return e;
  const Layout ly = make_layout(n, h, w_);
  if (int e = check_workspace(workspace, workspace_bytes, ly)) return e;
  int sms = 0;
  if (int e = check_arch(&sms)) return e;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return COATTN_E_DRIVER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tm_k, tm_v;
  const uint64_t t_rows = (uint64_t)2 * n * ly.Lp;
  if (int e = make_tmap(enc, &tm_k, seg(workspace, ly.off_t), t_rows, kC, kBN, bf16)) return e;
  if (int e = make_tmap(enc, &tm_v, seg(workspace, ly.off_vv), (uint64_t)2 * n * kC, ly.Lp, kC, bf16)) return e;
  if (!(flags & COATTN_FLAG_SINGLE_CTA)) {
    // default: CTA-pair kernel (cluster of 2, tcgen05 cta_group::2)
    CUtensorMap tm_q, tm_k2, tm_v2;
    if (int e = make_tmap(enc, &tm_q, seg(workspace, ly.off_t), t_rows, kC, k2BM, bf16)) return e;
    if (int e = make_tmap(enc, &tm_k2, seg(workspace, ly.off_t), t_rows, kC, k2BN / 2, bf16)) return e;
    if (int e = make_tmap(enc, &tm_v2, seg(workspace, ly.off_vv), (uint64_t)2 * n * kC, ly.Lp, kC / 2, bf16)) return e;
    Attend2Params q;
    q.z = z;
    q.lse = lse ? lse : reinterpret_cast<float*>(seg(workspace, ly.off_lse));
    q.cat_a = cat_a; q.cat_b = cat_b; q.mask = mask; q.gate_w = gate_w; q.gate_b = gate_b;
    q.v_a = v_a; q.v_b = v_b;
    q.N = n; q.L = ly.L; q.Lp = ly.Lp;
    q.q_pairs = (ly.L + 2 * k2BM - 1) / (2 * k2BM);
    q.kv_tiles = (ly.L + k2BN - 1) / k2BN;
    q.num_items = 2 * n * q.q_pairs;
    auto kern2 = bf16 ? attend2_kernel<true> : attend2_kernel<false>;
    cudaError_t e2 = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes);
    if (e2 != cudaSuccess) return (int)e2;
    int clusters = sms / 2;
    if (q.num_items < clusters) clusters = q.num_items;
    kern2<<<2 * clusters, k2Threads, k2SmemBytes, st>>>(tm_q, tm_k2, tm_v2, q);
    return (int)cudaGetLastError();
  }
  AttendParams p;
  p.t = reinterpret_cast<const unsigned short*>(seg(workspace, ly.off_t));
  p.z = z;
  p.lse = lse ? lse : reinterpret_cast<float*>(seg(workspace, ly.off_lse));
  p.cat_a = cat_a;
  p.cat_b = cat_b;
  p.mask = mask;
  p.gate_w = gate_w;
  p.gate_b = gate_b;
  p.N = n;
  p.L = ly.L;
  p.Lp = ly.Lp;
  p.q_tiles = (ly.L + kBM - 1) / kBM;
  p.kv_tiles = (ly.L + kBN - 1) / kBN;
  p.num_items = 2 * n * p.q_tiles;
  auto kern = bf16 ? attend_kernel<true> : attend_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttendSmemBytes);
  if (e != cudaSuccess) return (int)e;
  int grid = p.num_items < sms ? p.num_items : sms;
#ifdef COATTN_EXPERIMENT
  if (const char* g = getenv("COATTN_GRID")) { const int v = atoi(g); if (v > 0 && v < grid) grid = v; }
#endif
  p.trace = nullptr;
#ifdef COATTN_TRACE
  static long long* dtrace = nullptr;
  if (!dtrace) cudaMalloc(&dtrace, 96 * 8 * sizeof(long long));
  cudaMemsetAsync(dtrace, 0, 96 * 8 * sizeof(long long), st);
  p.trace = dtrace;
#endif
  kern<<<grid, kAttendThreads, kAttendSmemBytes, st>>>(tm_k, tm_v, p);
#ifdef COATTN_TRACE
  {
    static int calls = 0;
    if (++calls == 3) {
      long long hbuf[96 * 8];
      cudaStreamSynchronize(st);
      cudaMemcpy(hbuf, dtrace, sizeof(hbuf), cudaMemcpyDeviceToHost);
      const long long t0 = hbuf[0];
      for (int i = 0; i < 40; ++i)
        printf("tile %2d: loop+%6lld | S_issued +%5lld | p_full +%5lld | v_full +%5lld | PV_issued +%5lld || softmax saw S(j) at +%6lld\n", i,
               hbuf[i * 8 + 0] - t0, hbuf[i * 8 + 1] - hbuf[i * 8 + 0], hbuf[i * 8 + 2] - hbuf[i * 8 + 1],
               hbuf[i * 8 + 3] - hbuf[i * 8 + 2], hbuf[i * 8 + 4] - hbuf[i * 8 + 3], hbuf[i * 8 + 5] - t0);
    }
  }
#endif
  return (int)cudaGetLastError();
}

int coattn_stage_attend(float* z, float* lse, void* workspace, int64_t workspace_bytes, int n, int c, int h,
                        int w_, unsigned flags, void* stream) {
  if (int e = check_dims(n, c, h, w_)) return e;
  if (!workspace) return COATTN_E_NULL;
  float* zbuf = z ? z : reinterpret_cast<float*>(seg(workspace, make_layout(n, h, w_).off_z));
  return launch_attend(nullptr, nullptr, nullptr, nullptr, zbuf, lse, nullptr, nullptr, nullptr, workspace,
                       workspace_bytes, n, c, h, w_, flags, stream);
}

int coattn_stage_attend_gate(const float* v_a, const float* v_b, float* cat_a, float* cat_b, float* z, float* lse,
                             float* mask, const float* gate_w, const float* gate_b, void* workspace,
                             int64_t workspace_bytes, int n, int c, int h, int w_, unsigned flags, void* stream) {
  if (!cat_a || !cat_b || !gate_w) return COATTN_E_NULL;
  if ((v_a == nullptr) != (v_b == nullptr)) return COATTN_E_NULL;
  const bool pair = !(flags & COATTN_FLAG_SINGLE_CTA);
  if (int e = launch_attend(pair ? v_a : nullptr, pair ? v_b : nullptr, cat_a, cat_b, z, lse, mask, gate_w, gate_b,
                            workspace, workspace_bytes, n, c, h, w_, flags, stream))
    return e;
  if (!pair && v_a) return coattn_stage_passthrough(v_a, v_b, cat_a, cat_b, n, c, h, w_, stream);
  return COATTN_OK;
}

int coattn_stage_passthrough(const float* v_a, const float* v_b, float* cat_a, float* cat_b, int n, int c, int h,
                             int w_, void* stream) {
  if (!v_a || !v_b || !cat_a || !cat_b) return COATTN_E_NULL;
  if (int e = check_dims(n, c, h, w_)) return e;
  if (int e = check_arch(nullptr)) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PassParams p;
  p.v_a = v_a; p.v_b = v_b; p.cat_a = cat_a; p.cat_b = cat_b; p.N = n;
  p.plane = (size_t)kC * h * w_;
  const uintptr_t ptrs = reinterpret_cast<uintptr_t>(v_a) | reinterpret_cast<uintptr_t>(v_b) |
                         reinterpret_cast<uintptr_t>(cat_a) | reinterpret_cast<uintptr_t>(cat_b);
  // ~16 float4 per thread
  const size_t n4 = p.plane / 4;
  int gx = (int)((n4 + 256 * 16 - 1) / (256 * 16));
  if (gx < 1) gx = 1;
  if ((p.plane % 4) == 0 && (ptrs & 15) == 0) passthrough_kernel<4><<<dim3(gx, 2 * n), 256, 0, st>>>(p);
  else passthrough_kernel<1><<<dim3(gx, 2 * n), 256, 0, st>>>(p);
  return (int)cudaGetLastError();
}

int coattn_stage_gate(const float* z, const float* v_a, const float* v_b, const float* gate_w, const float* gate_b,
                      float* cat_a, float* cat_b, int n, int c, int h, int w_, void* stream) {
  if (!z || !v_a || !v_b || !gate_w || !cat_a || !cat_b) return COATTN_E_NULL;
  if (int e = check_dims(n, c, h, w_)) return e;
  if (int e = check_arch(nullptr)) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GateParams p;
  p.z = z; p.v_a = v_a; p.v_b = v_b; p.gate_w = gate_w; p.gate_b = gate_b;
  p.cat_a = cat_a; p.cat_b = cat_b; p.N = n; p.L = h * w_;
  const uintptr_t ptrs = reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(v_a) |
                         reinterpret_cast<uintptr_t>(v_b) | reinterpret_cast<uintptr_t>(cat_a) |
                         reinterpret_cast<uintptr_t>(cat_b);
  if ((p.L % 4) == 0 && (ptrs & 15) == 0) {
    gate_kernel<4><<<dim3((p.L + 127) / 128, 2 * n), kGateThreads, 0, st>>>(p);
  } else {
    gate_kernel<1><<<dim3((p.L + 31) / 32, 2 * n), kGateThreads, 0, st>>>(p);
  }
  return (int)cudaGetLastError();
}

int coattn_forward(const float* v_a, const float* v_b, const float* w, const float* gate_w, const float* gate_b,
                   float* cat_a, float* cat_b, float* z, float* lse, float* mask, void* workspace,
                   int64_t workspace_bytes, int n, int c, int h, int w_, unsigned flags, void* stream) {
  if (!v_a || !v_b || !w || !gate_w || !cat_a || !cat_b) return COATTN_E_NULL;
  if (int e = check_dims(n, c, h, w_)) return e;
  const Layout ly = make_layout(n, h, w_);
  if (int e = check_workspace(workspace, workspace_bytes, ly)) return e;
  if (int e = coattn_stage_prep(v_a, v_b, w, workspace, workspace_bytes, n, c, h, w_, flags, stream)) return e;
  if (int e = coattn_stage_project(workspace, workspace_bytes, n, c, h, w_, flags, stream)) return e;
  if (flags & COATTN_FLAG_UNFUSED_GATE) {
    float* zbuf = z ? z : reinterpret_cast<float*>(seg(workspace, ly.off_z));
    if (int e = coattn_stage_attend(zbuf, lse, workspace, workspace_bytes, n, c, h, w_, flags, stream)) return e;
    return coattn_stage_gate(zbuf, v_a, v_b, gate_w, gate_b, cat_a, cat_b, n, c, h, w_, stream);
  }
  return coattn_stage_attend_gate(v_a, v_b, cat_a, cat_b, z, lse, mask, gate_w, gate_b, workspace, workspace_bytes, n, c,
                                  h, w_, flags, stream);
}

}  // extern "C"
