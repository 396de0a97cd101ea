// Host side of the C ABI declared in include/coattn_b200.h: argument checks, workspace carving,
// TMA descriptor encoding and kernel launches.  No torch types, no global mutable state.
#include "../../include/coattn_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <cuda.h>
#include <cuda_runtime.h>

#include <utility>
#include "coattn_kernels.cuh"
#include "attend2_kernel.cuh"
#include "backward_kernels.cuh"
#include "bwd_flash_kernel.cuh"

namespace {

using namespace coattn;

constexpr int64_t kAlign = 1024;
constexpr unsigned kInternalPrepOnlyB = 1u << 30;   // internal: prep converts V_b only (V_a goes through project_fused)
constexpr unsigned kInternalNeedQ16 = 1u << 29;     // internal: the caller (backward) needs the projected plane Q16 in the workspace
constexpr unsigned kInternalW16Ready = 1u << 28;    // internal: W16 of the workspace is already written (bwd_init_kernel)
inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Layout {
  int N, L, Lp;
  int parts;   // key-range parts the z / lse segments have room for (COATTN_FLAG_SPLIT_KEYS)
  int64_t off_status, off_t, off_at, off_vv, off_w16, off_z, off_lse, total;
  int64_t bytes_t, bytes_at, bytes_vv, bytes_w16, bytes_z, bytes_lse;
  // element strides
  int64_t t_pass_elems() const { return (int64_t)N * Lp * kC; }   // T[pass] and VV[pass]
};

// COATTN_FLAG_SPLIT_KEYS: parts per item.  `items` work items of T key tiles each on `clusters` CTA pairs: split only
// when at least two parts per item fit beside each other, never below four key tiles per part.
int choose_splits(int items, int T, int clusters) {
  if (items < 1 || 2 * items > clusters) return 1;
  int s = clusters / items;
  if (s > kMaxKeySplits) s = kMaxKeySplits;
  if (s > T / 4) s = T / 4;
  return s < 1 ? 1 : s;
}

Layout make_layout(int n, int h, int w) {
  Layout ly{};
  ly.N = n;
  ly.L = h * w;
  ly.Lp = (int)round_up(ly.L, kLPad);
  const int64_t plane = (int64_t)n * ly.Lp * kC * 2;  // one bf16 [N][Lp][C] (or [N][C][Lp]) array
  int64_t off = 0;
  ly.off_status = off; off = kAlign;   // status block (coattn_status_*): always the first bytes, whatever the problem size
  ly.off_t = off;   ly.bytes_t = 2 * plane;   off = round_up(off + ly.bytes_t, kAlign);
  ly.off_at = off;  ly.bytes_at = plane;      off = round_up(off + ly.bytes_at, kAlign);
  ly.off_vv = off;  ly.bytes_vv = 3 * plane;  off = round_up(off + ly.bytes_vv, kAlign);   // B16, A16, Q16 ([C][Lp] each)
  ly.off_w16 = off; ly.bytes_w16 = (int64_t)kC * kC * 2; off = round_up(off + ly.bytes_w16, kAlign);
  // room for the parts of COATTN_FLAG_SPLIT_KEYS (more than one only for the few-pair shapes that can use them:
  // frame-A-only items on the 74 CTA pairs of a B200 is the case with the most parts)
  const int parts = choose_splits(n * ((ly.L + 2 * k2BM - 1) / (2 * k2BM)), (ly.L + k2BN - 1) / k2BN, 74);
  ly.off_z = off;   ly.bytes_z = (int64_t)parts * 2 * n * kC * ly.L * 4; off = round_up(off + ly.bytes_z, kAlign);
  ly.off_lse = off; ly.bytes_lse = (int64_t)parts * 2 * n * ly.L * 4;    off = round_up(off + ly.bytes_lse, kAlign);
  ly.parts = parts;
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

// Immutable per-device attributes, queried once per device and process (the only state the library keeps).
// Written with relaxed atomics: every thread that races here writes the same values.
constexpr int kMaxDevices = 64;
std::atomic<int> g_dev_major[kMaxDevices];
std::atomic<int> g_dev_sms[kMaxDevices];

int check_arch(int* sm_count) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  int major = 0, sms = 0;
  if (dev >= 0 && dev < kMaxDevices && (major = g_dev_major[dev].load(std::memory_order_relaxed)) != 0) {
    sms = g_dev_sms[dev].load(std::memory_order_relaxed);
  } else {
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < kMaxDevices) {
      g_dev_sms[dev].store(sms, std::memory_order_relaxed);
      g_dev_major[dev].store(major, std::memory_order_relaxed);
    }
  }
  if (major != 10) return COATTN_E_ARCH;
  if (sm_count) *sm_count = sms;
  return COATTN_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn resolve_encode_fn() {
  // resolved through the runtime so the library has no link-time dependency on libcuda
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess) return nullptr;
  if (qres != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<EncodeTiledFn>(fn);
}
EncodeTiledFn get_encode_fn() {
  static std::atomic<EncodeTiledFn> cached{nullptr};      // the entry point does not change during the life of the process
  EncodeTiledFn fn = cached.load(std::memory_order_relaxed);
  if (!fn) {
    fn = resolve_encode_fn();
    if (fn) cached.store(fn, std::memory_order_relaxed);
  }
  return fn;
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

// Launch with programmatic stream serialization: the grid may be scheduled while its predecessor in the stream drains; the
// kernel itself waits (pdl_wait, first statement) until the predecessor has completed.  Only for kernels that do so.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

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
    case COATTN_E_UNSUPPORTED: return "valid request that is not implemented (counterpart-frame gradients, or a flag combination)";
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
  else if (!strcmp(name, "status")) { *offset = ly.off_status; *bytes = COATTN_STATUS_WORDS * 4; }
  else return COATTN_E_NULL;
  return COATTN_OK;
}

int coattn_status_clear(void* workspace, void* stream) {
  if (!workspace) return COATTN_E_NULL;
  return (int)cudaMemsetAsync(workspace, 0, COATTN_STATUS_WORDS * 4, static_cast<cudaStream_t>(stream));
}

int coattn_status_read(const void* workspace, uint32_t* host_words, void* stream) {
  if (!workspace || !host_words) return COATTN_E_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemcpyAsync(host_words, workspace, COATTN_STATUS_WORDS * 4, cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaStreamSynchronize(st);
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
  p.only_b = (flags & kInternalPrepOnlyB) ? 1 : 0;
  unsigned short* w16 = reinterpret_cast<unsigned short*>(seg(workspace, ly.off_w16));
  const dim3 grid(ly.Lp / kPrepTileL, p.only_b ? n : 2 * n);
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

// prep (V_b only) + fused A-side prep/projection: the default forward path
static int prep_and_project_fused(const float* v_a, const float* v_b, const float* w, void* workspace,
                                  int64_t workspace_bytes, int n, int c, int h, int w_, unsigned flags, void* stream) {
  if (int e = coattn_stage_prep(v_a, v_b, w, workspace, workspace_bytes, n, c, h, w_, flags | kInternalPrepOnlyB, stream))
    return e;
  const bool bf16 = (flags & COATTN_FLAG_BF16) != 0;
  const Layout ly = make_layout(n, h, w_);
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return COATTN_E_DRIVER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tm_w;
  if (int e = make_tmap(enc, &tm_w, seg(workspace, ly.off_w16), kC, kC, 256, bf16)) return e;
  ProjectFusedParams p;
  p.va = v_a;
  p.a16 = reinterpret_cast<unsigned short*>(seg(workspace, ly.off_vv)) + ly.t_pass_elems();
  p.qt = reinterpret_cast<unsigned short*>(seg(workspace, ly.off_t)) + ly.t_pass_elems();
  p.L = ly.L;
  p.Lp = ly.Lp;
  auto kern = bf16 ? project_fused_kernel<true> : project_fused_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kProjFusedSmemBytes);
  if (e != cudaSuccess) return (int)e;
  kern<<<dim3(ly.Lp / 128, n), kNumThreads, kProjFusedSmemBytes, st>>>(tm_w, p);
  return (int)cudaGetLastError();
}

// MN-major path: cast both frames (no transposes) + channel-major projection.
//   x   [3][N][C][Lp] 16-bit planes: 0 = V_b, 1 = V_a (written by the cast), 2 = Q = W V_a (written by the projection)
//   w16 [C][C] 16-bit copy of W
//   in16: 0 = fp32 features (cast), 1 = 16-bit features copied into the padded planes (pad16_kernel),
//         2 = 16-bit features consumed in place: nothing is copied and the projection reads V_a through its own map
static int cast_project_core(const void* v_a, const void* v_b, const float* w, unsigned short* x, unsigned short* w16,
                             int n, const Layout& ly, bool bf16, int project, cudaStream_t st, int n_a = -1,
                             int in16 = 0, unsigned* status = nullptr, bool w_ready = false) {
  // project: 0 = cast only, 1 = cast + W16 + Q16 = W V_a (project_mn), 2 = cast + W16 (the attend kernel projects itself)
  if (n_a < 0) n_a = n;       // samples of V_a (query frames); the planes are laid out for n samples either way
  CastParams cp;
  cp.va = static_cast<const float*>(v_a);   // fp32 features; with in16 == 1 pad16_kernel reads the same pointers as 16-bit data
  cp.vb = static_cast<const float*>(v_b);
  cp.x = x; cp.N = n; cp.L = ly.L; cp.Lp = ly.Lp; cp.Na = n_a;
  cp.status = status;
  cp.first_plane = 0;
  // W is cast by the first C blocks of the feature cast (256 threads = one row each); a launch of its own only when the
  // features need no cast kernel (16-bit interfaces, planes written by the fused encoder tail)
  const bool w_in_cast = project != 0 && in16 == 0 && !w_ready;
  cp.w = w_in_cast ? w : nullptr;
  cp.w16 = w16;
  const bool vec = (ly.L % 4 == 0) && (((reinterpret_cast<uintptr_t>(v_a) | reinterpret_cast<uintptr_t>(v_b)) & 15) == 0);
  const dim3 cgrid(n * kC, 2);
  if (in16 == 1) {
    if (((reinterpret_cast<uintptr_t>(v_a) | reinterpret_cast<uintptr_t>(v_b)) & 3) == 0) pad16_kernel<true><<<cgrid, 256, 0, st>>>(cp);
    else pad16_kernel<false><<<cgrid, 256, 0, st>>>(cp);
  } else if (in16 == 2 || in16 == 3) {
    // 2: 16-bit features consumed in place; 3: the planes were already written by coattn_stage_tail
  } else if (bf16) {
    if (cudaError_t e = launch_pdl(vec ? cast_kernel<true, 4> : cast_kernel<true, 1>, cgrid, dim3(256), 0, st, cp)) return (int)e;
  } else {
    if (cudaError_t e = launch_pdl(vec ? cast_kernel<false, 4> : cast_kernel<false, 1>, cgrid, dim3(256), 0, st, cp)) return (int)e;
  }
  if (!project) return (int)cudaGetLastError();
  if (!w_in_cast && !w_ready) {
    if (bf16) cast_w_kernel<true><<<(kC * kC + 255) / 256, 256, 0, st>>>(w, w16, kC * kC);
    else cast_w_kernel<false><<<(kC * kC + 255) / 256, 256, 0, st>>>(w, w16, kC * kC);
  }
  if (project == 2) return (int)cudaGetLastError();
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return COATTN_E_DRIVER;
  CUtensorMap tm_w, tm_x;
  if (int e = make_tmap(enc, &tm_w, w16, kC, kC, 128, bf16)) return e;
  if (in16 == 2) {     // positions past L are out of bounds of the map: TMA fills them with zeros, as the padded planes do
    if (int e = make_tmap(enc, &tm_x, v_a, (uint64_t)n_a * kC, ly.L, 256, bf16)) return e;
  } else {
    if (int e = make_tmap(enc, &tm_x, x, (uint64_t)3 * n * kC, ly.Lp, 256, bf16)) return e;
  }
  int sms = 148;
  if (int e = check_arch(&sms)) return e;
  ProjectMnParams pp;
  pp.q16 = x + 2 * ly.t_pass_elems();
  pp.Lp = ly.Lp;
  pp.tiles_per_sample = ly.Lp / kProjMnTile;
  pp.num_tiles = n_a * pp.tiles_per_sample;
  pp.a_row0_base = (in16 == 2) ? 0 : n * kC;
  pp.status = status;
  auto kern = bf16 ? project_mn_kernel<true> : project_mn_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kProjMnSmemBytes);
  if (e != cudaSuccess) return (int)e;
  if ((e = launch_pdl(kern, dim3(pp.num_tiles < sms ? pp.num_tiles : sms), dim3(kProjMnThreads), kProjMnSmemBytes, st, tm_w, tm_x, pp)) !=
      cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}

// the attend kernel projects its query tiles itself (FOLD) on the default channel-major path
static inline bool uses_fold(unsigned flags) {
  return !(flags & (COATTN_FLAG_UNFOLDED | COATTN_FLAG_KMAJOR | COATTN_FLAG_UNFUSED_PREP | COATTN_FLAG_SINGLE_CTA | COATTN_FLAG_SOFTMAX16));
}

static int cast_and_project_mn(const float* v_a, const float* v_b, const float* w, void* workspace, int64_t workspace_bytes,
                               int n, int c, int h, int w_, unsigned flags, void* stream) {
  if (int e = check_dims(n, c, h, w_)) return e;
  const Layout ly = make_layout(n, h, w_);
  if (int e = check_workspace(workspace, workspace_bytes, ly)) return e;
  if (int e = check_arch(nullptr)) return e;
  return cast_project_core(v_a, v_b, w, reinterpret_cast<unsigned short*>(seg(workspace, ly.off_vv)),
                           reinterpret_cast<unsigned short*>(seg(workspace, ly.off_w16)), n, ly,
                           (flags & COATTN_FLAG_BF16) != 0, (uses_fold(flags) && !(flags & kInternalNeedQ16)) ? 2 : 1,
                           static_cast<cudaStream_t>(stream), -1, (flags & COATTN_FLAG_PLANES_READY) ? 3 : 0,
                           reinterpret_cast<unsigned*>(seg(workspace, ly.off_status)), (flags & kInternalW16Ready) != 0);
}

extern "C" int coattn_stage_tail(const float* x, const float* scale, const float* shift, const float* slope, float* y,
                                 void* workspace, int64_t workspace_bytes, int frame, int n, int c, int h, int w_,
                                 unsigned flags, void* stream) {
  if (!x || !scale || !shift || !slope) return COATTN_E_NULL;
  if (frame != 0 && frame != 1) return COATTN_E_SHAPE;
  if (int e = check_dims(n, c, h, w_)) return e;
  const Layout ly = make_layout(n, h, w_);
  if (int e = check_workspace(workspace, workspace_bytes, ly)) return e;
  if (int e = check_arch(nullptr)) return e;
  TailParams p;
  p.x = x; p.scale = scale; p.shift = shift; p.slope = slope; p.y = y;
  // planes X = [B16, A16, Q16]: frame A (0) -> plane 1, frame B (1) -> plane 0
  p.plane = reinterpret_cast<unsigned short*>(seg(workspace, ly.off_vv)) + (frame == 0 ? ly.t_pass_elems() : 0);
  p.L = ly.L; p.Lp = ly.Lp;
  p.status = reinterpret_cast<unsigned*>(seg(workspace, ly.off_status));
  p.status_plane = frame == 0 ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = (ly.L % 4 == 0) && (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0);
  const bool bf16 = (flags & COATTN_FLAG_BF16) != 0;
  const dim3 grid(n * kC);
  if (bf16) { if (vec) aspp_tail_kernel<true, 4><<<grid, 256, 0, st>>>(p); else aspp_tail_kernel<true, 1><<<grid, 256, 0, st>>>(p); }
  else      { if (vec) aspp_tail_kernel<false, 4><<<grid, 256, 0, st>>>(p); else aspp_tail_kernel<false, 1><<<grid, 256, 0, st>>>(p); }
  return (int)cudaGetLastError();
}

extern "C" int coattn_stage_prep_project(const float* v_a, const float* v_b, const float* w, void* workspace,
                                        int64_t workspace_bytes, int n, int c, int h, int w_, unsigned flags,
                                        void* stream) {
  if (!v_a || !v_b || !w) return COATTN_E_NULL;
  if (int e = check_dims(n, c, h, w_)) return e;
  if (!(flags & (COATTN_FLAG_KMAJOR | COATTN_FLAG_SINGLE_CTA | COATTN_FLAG_UNFUSED_PREP)))
    return cast_and_project_mn(v_a, v_b, w, workspace, workspace_bytes, n, c, h, w_, flags, stream);
  if (((h * w_) % 2 != 0) || (reinterpret_cast<uintptr_t>(v_a) & 7) != 0) {   // odd L: separate kernels
    if (int e = coattn_stage_prep(v_a, v_b, w, workspace, workspace_bytes, n, c, h, w_, flags, stream)) return e;
    return coattn_stage_project(workspace, workspace_bytes, n, c, h, w_, flags, stream);
  }
  return prep_and_project_fused(v_a, v_b, w, workspace, workspace_bytes, n, c, h, w_, flags, stream);
}

static int launch_attend(const void* v_a, const void* v_b, void* cat_a, void* cat_b, float* z, float* lse,
                         float* mask, const float* gate_w, const float* gate_b, void* workspace, int64_t workspace_bytes, int n, int c, int h, int w_,
                         unsigned flags, void* stream, int q_group = 1, int in16 = 0, int splits = 1) {
  // in16 (coattn_forward16): v_a, v_b, cat_a, cat_b hold 16-bit elements; 2 = the operands are read from v_a / v_b in place
  const bool bf16 = (flags & COATTN_FLAG_BF16) != 0;
  if (int e = check_dims(n, c, h, w_)) return e;
  const Layout ly = make_layout(n, h, w_);
  if (int e = check_workspace(workspace, workspace_bytes, ly)) return e;
  int sms = 0;
  if (int e = check_arch(&sms)) return e;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return COATTN_E_DRIVER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t t_rows = (uint64_t)2 * n * ly.Lp;
  if (!(flags & COATTN_FLAG_SINGLE_CTA)) {
    // default: CTA-pair kernel (cluster of 2, tcgen05 cta_group::2)
    const bool mn = !(flags & (COATTN_FLAG_KMAJOR | COATTN_FLAG_UNFUSED_PREP));
    CUtensorMap tm_q, tm_k2, tm_v2, tm_v3, tm_w0, tm_w1;
    Attend2Params q;
    if (in16 && !mn) return COATTN_E_UNSUPPORTED;
    const bool fold = mn && uses_fold(flags);
    if (fold) {
      if (int e = make_tmap(enc, &tm_w0, seg(workspace, ly.off_w16), kC, kC, 64, bf16)) return e;
      if (int e = make_tmap(enc, &tm_w1, seg(workspace, ly.off_w16), kC, kC, 256, bf16)) return e;
    } else {
      memset(&tm_w0, 0, sizeof(tm_w0));
      memset(&tm_w1, 0, sizeof(tm_w1));
    }
    if (in16 == 2) {
      // 16-bit features consumed in place: one map per tensor (positions past L are zero-filled by TMA); Q16 is the
      // only operand that lives in the workspace
      const int n_a = n / q_group;
      const uint8_t* q16 = seg(workspace, ly.off_vv) + 2 * ly.t_pass_elems() * 2;
      if (fold) {      // queries are projected in the kernel from V_a itself
        if (int e = make_tmap(enc, &tm_q, v_a, (uint64_t)n_a * kC, ly.L, kC, bf16)) return e;
      } else if (int e = make_tmap(enc, &tm_q, q16, (uint64_t)n * kC, ly.Lp, kC, bf16)) return e;
      if (int e = make_tmap(enc, &tm_k2, v_b, (uint64_t)n * kC, ly.L, kC, bf16)) return e;
      if (int e = make_tmap(enc, &tm_v2, v_b, (uint64_t)n * kC, ly.L, kC / 2, bf16)) return e;
      if (int e = make_tmap(enc, &tm_v3, v_a, (uint64_t)n_a * kC, ly.L, kC / 2, bf16)) return e;
      q.xq_row0 = 0; q.xb_row0 = 0; q.v0_row0 = 0; q.v1_row0 = 0;
    } else {
      if (mn) {
        // queries and keys straight from the channel-major planes X = [B16, A16, Q16]
        if (int e = make_tmap(enc, &tm_q, seg(workspace, ly.off_vv), (uint64_t)3 * n * kC, ly.Lp, kC, bf16)) return e;
        tm_k2 = tm_q;
      } else {
        if (int e = make_tmap(enc, &tm_q, seg(workspace, ly.off_t), t_rows, kC, k2BM, bf16)) return e;
        if (int e = make_tmap(enc, &tm_k2, seg(workspace, ly.off_t), t_rows, kC, k2BN / 2, bf16)) return e;
      }
      if (int e = make_tmap(enc, &tm_v2, seg(workspace, ly.off_vv), (uint64_t)3 * n * kC, ly.Lp, kC / 2, bf16)) return e;
      tm_v3 = tm_v2;
      q.xq_row0 = (fold ? 1 : 2) * n * kC; q.xb_row0 = 0; q.v0_row0 = 0; q.v1_row0 = n * kC;      // fold: "queries" = plane A16
    }
    q.status = reinterpret_cast<unsigned*>(seg(workspace, ly.off_status));
    q.z = z;
    q.lse = lse ? lse : reinterpret_cast<float*>(seg(workspace, ly.off_lse));
    q.cat_a = cat_a; q.cat_b = cat_b; q.mask = mask; q.gate_w = gate_w; q.gate_b = gate_b;
    q.v_a = (flags & COATTN_FLAG_GATED_ONLY) ? nullptr : v_a;
    q.v_b = (flags & COATTN_FLAG_GATED_ONLY) ? nullptr : v_b;
    q.out_channels = (flags & COATTN_FLAG_GATED_ONLY) ? kC : 2 * kC;
    q.N = n; q.L = ly.L; q.Lp = ly.Lp;
    q.q_pairs = (ly.L + 2 * k2BM - 1) / (2 * k2BM);
    q.kv_tiles = (ly.L + k2BN - 1) / k2BN;
    q.passes = (flags & COATTN_FLAG_A_ONLY) ? 1 : 2;
    q.q_group = q_group;
    q.splits = splits;
    if (q_group != 1 && (q.passes != 1 || !mn)) return COATTN_E_UNSUPPORTED;
    if (splits != 1 && (cat_a || v_a || mask || !z || !lse || !mn)) return COATTN_E_UNSUPPORTED;   // parts: z / lse only
    q.num_items = q.passes * n * q.q_pairs;
    // 8 softmax warps (two column groups per TMEM lane quadrant) by default; COATTN_FLAG_SOFTMAX16 selects the 16-warp
    // layout (four groups)
    const bool g4 = (flags & COATTN_FLAG_SOFTMAX16) != 0;
    void (*kern2)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, Attend2Params);
    if (in16) {
      if (g4) return COATTN_E_UNSUPPORTED;
      if (fold) kern2 = bf16 ? attend2_kernel<true, true, 2, true, false, true> : attend2_kernel<false, true, 2, true, false, true>;
      else kern2 = bf16 ? attend2_kernel<true, true, 2, true> : attend2_kernel<false, true, 2, true>;
    } else if (splits != 1) {
      if (g4) return COATTN_E_UNSUPPORTED;
      if (fold) kern2 = bf16 ? attend2_kernel<true, true, 2, false, true, true> : attend2_kernel<false, true, 2, false, true, true>;
      else kern2 = bf16 ? attend2_kernel<true, true, 2, false, true> : attend2_kernel<false, true, 2, false, true>;
    } else if (fold) {
      kern2 = bf16 ? attend2_kernel<true, true, 2, false, false, true> : attend2_kernel<false, true, 2, false, false, true>;
    } else if (g4) kern2 = mn ? (bf16 ? attend2_kernel<true, true, 4> : attend2_kernel<false, true, 4>)
                       : (bf16 ? attend2_kernel<true, false, 4> : attend2_kernel<false, false, 4>);
    else    kern2 = mn ? (bf16 ? attend2_kernel<true, true, 2> : attend2_kernel<false, true, 2>)
                       : (bf16 ? attend2_kernel<true, false, 2> : attend2_kernel<false, false, 2>);
    const int smem2 = g4 ? Attend2Cfg<4>::kSmemBytes : Attend2Cfg<2>::kSmemBytes;
    const int threads2 = g4 ? Attend2Cfg<4>::kThreads : Attend2Cfg<2>::kThreads;
    cudaError_t e2 = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    if (e2 != cudaSuccess) return (int)e2;
    int clusters = sms / 2;
    if (q.num_items * splits < clusters) clusters = q.num_items * splits;
    if ((e2 = launch_pdl(kern2, dim3(2 * clusters), dim3(threads2), smem2, st, tm_q, tm_k2, tm_v2, tm_v3, tm_w0, tm_w1, q)) != cudaSuccess)
      return (int)e2;
#ifdef COATTN_TRACE2
    {
      static int calls = 0;
      if (++calls == 3) {
        long long hb[16 * 16];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(hb, g_attend2_trace, sizeof(hb));
        long long tb[32 * 8];
        cudaMemcpyFromSymbol(tb, g_attend2_tiles, sizeof(tb));
        for (int j = 1; j < 29; ++j) {
          const long long* r = tb + j * 8;
          printf("tile %2d: MMA loop top +%6lld | S(j+2): waits+issue %5lld | then P(j) seen after %5lld | PV issue %4lld"
                 " || softmax saw S at +%6lld | busy %5lld | period %5lld | P(j) arrive (warp 0) +%6lld\n", j, r[2] - tb[8 + 2],
                 r[3] - r[2], r[0] - r[3], r[1] - r[0], r[4] - tb[8 + 2], r[5] - r[4], r[4] - tb[(j - 1) * 8 + 4], r[5] - tb[8 + 2]);
        }
        long long wb[2 * 32 * 16];
        cudaMemcpyFromSymbol(wb, g_attend2_warps, sizeof(wb));
        for (int j = 1; j < 29; ++j) {
          printf("tile %2d warps 0-7 saw S / arrived P (relative to warp 0 seeing S):", j);
          for (int w = 0; w < 8; ++w) printf("  %5lld/%5lld", wb[j * 16 + w] - wb[j * 16], wb[512 + j * 16 + w] - wb[j * 16]);
          printf("\n");
        }
        const long long t0 = hb[0];
        for (int it = 0; it < 14; ++it) {
          const long long* r = hb + it * 16;
          printf("item %2d  MMA: start +%7lld | q_full +%5lld | S(0), S(1) issued +%5lld | P(0) ready +%5lld | last PV issued +%6lld"
                 "  || softmax: start +%7lld | S(0) ready +%5lld | last P +%6lld | O complete +%5lld | l xchg +%5lld | first ld +%5lld"
                 " | dot pass +%5lld | xchg +%5lld | sigmoid+first ld +%5lld | stores +%5lld\n", it,
                 r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[8] - t0, r[9] - r[8], r[10] - r[9],
                 r[11] - r[10], r[5] - r[11], r[6] - r[5], r[13] - r[6], r[14] - r[13], r[7] - r[14], r[12] - r[7]);
        }
      }
    }
#endif
    return (int)cudaGetLastError();
  }
  if (flags & (COATTN_FLAG_A_ONLY | COATTN_FLAG_GATED_ONLY)) return COATTN_E_UNSUPPORTED;   // cross-check kernel: full concat only
  CUtensorMap tm_k, tm_v;
  if (int e = make_tmap(enc, &tm_k, seg(workspace, ly.off_t), t_rows, kC, kBN, bf16)) return e;
  if (int e = make_tmap(enc, &tm_v, seg(workspace, ly.off_vv), (uint64_t)2 * n * kC, ly.Lp, kC, bf16)) return e;
  AttendParams p;
  p.t = reinterpret_cast<const unsigned short*>(seg(workspace, ly.off_t));
  p.z = z;
  p.lse = lse ? lse : reinterpret_cast<float*>(seg(workspace, ly.off_lse));
  p.cat_a = static_cast<float*>(cat_a);     // the cross-check kernel is fp32 only (coattn_forward16 rejects its flag)
  p.cat_b = static_cast<float*>(cat_b);
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
  if (!cat_a || !gate_w) return COATTN_E_NULL;
  if (!cat_b && !(flags & COATTN_FLAG_A_ONLY)) return COATTN_E_NULL;
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
  if ((p.L % 2) == 0 && (ptrs & 7) == 0) {
    // [256 ch x 64 pos] tiles, float2 per lane, two blocks per SM: 6.2 TB/s at batch 32 (0.95 of the measured copy peak);
    // the one-block-per-SM float4 form (-DCOATTN_GATE_VEC4, [256 x 128] tiles) reaches 6.05
#ifdef COATTN_GATE_VEC4
    if ((p.L % 4) == 0 && (ptrs & 15) == 0) gate_kernel<4><<<dim3((p.L + 127) / 128, 2 * n), kGateThreads, 0, st>>>(p);
    else
#endif
    gate_kernel<2, 2, 8><<<dim3((p.L + 63) / 64, 2 * n), kGateThreads, 0, st>>>(p);
  } else {
    gate_kernel<1><<<dim3((p.L + 31) / 32, 2 * n), kGateThreads, 0, st>>>(p);
  }
  return (int)cudaGetLastError();
}

// COATTN_FLAG_SPLIT_KEYS: parts to use for this call (1 = the default single-sweep path)
static int split_parts(unsigned flags, int n, const Layout& ly, int* parts) {
  *parts = 1;
  if (!(flags & COATTN_FLAG_SPLIT_KEYS)) return COATTN_OK;
  if (flags & (COATTN_FLAG_UNFUSED_GATE | COATTN_FLAG_SINGLE_CTA | COATTN_FLAG_KMAJOR | COATTN_FLAG_UNFUSED_PREP |
               COATTN_FLAG_SOFTMAX16))
    return COATTN_E_UNSUPPORTED;
  int sms = 0;
  if (int e = check_arch(&sms)) return e;
  const int passes = (flags & COATTN_FLAG_A_ONLY) ? 1 : 2;
  int s = choose_splits(passes * n * ((ly.L + 2 * k2BM - 1) / (2 * k2BM)), (ly.L + k2BN - 1) / k2BN, sms / 2);
  if (s > ly.parts) s = ly.parts;
  *parts = s;
  return COATTN_OK;
}

// attend2 over `parts` key ranges per item (z / lse of every part in the workspace), then merge + gate + concat
static int attend_split_merge(const float* v_a, const float* v_b, const float* gate_w, const float* gate_b, float* cat_a,
                              float* cat_b, float* z, float* lse, float* mask, void* workspace, int64_t workspace_bytes,
                              int n, int c, int h, int w_, unsigned flags, void* stream, int q_group, int parts) {
  const Layout ly = make_layout(n, h, w_);
  float* zp = reinterpret_cast<float*>(seg(workspace, ly.off_z));
  float* lsep = reinterpret_cast<float*>(seg(workspace, ly.off_lse));
  if (int e = launch_attend(nullptr, nullptr, nullptr, nullptr, zp, lsep, nullptr, gate_w, gate_b, workspace, workspace_bytes,
                            n, c, h, w_, flags, stream, q_group, 0, parts))
    return e;
  MergeParams m;
  m.zp = zp; m.lsep = lsep;
  m.v_a = (flags & COATTN_FLAG_GATED_ONLY) ? nullptr : v_a;
  m.v_b = (flags & COATTN_FLAG_GATED_ONLY) ? nullptr : v_b;
  m.gate_w = gate_w; m.gate_b = gate_b; m.cat_a = cat_a; m.cat_b = cat_b; m.z = z; m.lse = lse; m.mask = mask;
  m.N = n; m.L = ly.L; m.splits = parts; m.passes = (flags & COATTN_FLAG_A_ONLY) ? 1 : 2;
  m.out_channels = (flags & COATTN_FLAG_GATED_ONLY) ? kC : 2 * kC;
  m.q_group = q_group;
  const uintptr_t ptrs = reinterpret_cast<uintptr_t>(v_a) | reinterpret_cast<uintptr_t>(v_b) |
                         reinterpret_cast<uintptr_t>(cat_a) | reinterpret_cast<uintptr_t>(cat_b) |
                         reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(lse) | reinterpret_cast<uintptr_t>(mask);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // few pairs by construction: 32-position blocks (4-byte accesses, 128-byte rows per warp) unless the 128-position
  // blocks of the float4 variant already give every SM two CTAs
  const bool vec4 = (ly.L % 4) == 0 && (ptrs & 15) == 0 && ((ly.L + 127) / 128) * m.passes * n >= 296;
  if (vec4) merge_gate_kernel<4><<<dim3((ly.L + 127) / 128, m.passes * n), kGateThreads, 0, st>>>(m);
  else merge_gate_kernel<1><<<dim3((ly.L + 31) / 32, m.passes * n), kGateThreads, 0, st>>>(m);
  return (int)cudaGetLastError();
}

int coattn_forward_queries(const float* v_a, const float* v_b, const float* w, const float* gate_w, const float* gate_b,
                           float* cat_a, void* workspace, int64_t workspace_bytes, int nq, int refs, int c, int h, int w_,
                           unsigned flags, void* stream) {
  if (!v_a || !v_b || !w || !gate_w || !cat_a) return COATTN_E_NULL;
  if (nq < 1 || refs < 1) return COATTN_E_SHAPE;
  if (flags & (COATTN_FLAG_UNFUSED_GATE | COATTN_FLAG_SINGLE_CTA | COATTN_FLAG_KMAJOR | COATTN_FLAG_UNFUSED_PREP))
    return COATTN_E_UNSUPPORTED;
  const int n = nq * refs;
  if (int e = check_dims(n, c, h, w_)) return e;
  const Layout ly = make_layout(n, h, w_);
  if (int e = check_workspace(workspace, workspace_bytes, ly)) return e;
  if (int e = check_arch(nullptr)) return e;
  // the query side (16-bit cast of V_a and Q = W V_a) is prepared once per query frame, not once per pair
  if (int e = cast_project_core(v_a, v_b, w, reinterpret_cast<unsigned short*>(seg(workspace, ly.off_vv)),
                                reinterpret_cast<unsigned short*>(seg(workspace, ly.off_w16)), n, ly,
                                (flags & COATTN_FLAG_BF16) != 0, uses_fold(flags) ? 2 : 1, static_cast<cudaStream_t>(stream), nq, 0,
                                reinterpret_cast<unsigned*>(seg(workspace, ly.off_status))))
    return e;
  int parts = 1;
  if (int e = split_parts(flags | COATTN_FLAG_A_ONLY, n, ly, &parts)) return e;
  if (parts > 1)
    return attend_split_merge(v_a, v_b, gate_w, gate_b, cat_a, nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes,
                              n, c, h, w_, (flags | COATTN_FLAG_A_ONLY) & ~COATTN_FLAG_SPLIT_KEYS, stream, refs, parts);
  return launch_attend(v_a, v_b, cat_a, nullptr, nullptr, nullptr, nullptr, gate_w, gate_b, workspace, workspace_bytes, n, c,
                       h, w_, (flags | COATTN_FLAG_A_ONLY) & ~COATTN_FLAG_SPLIT_KEYS, stream, refs);
}

int coattn_forward16(const void* v_a, const void* v_b, const float* w, const float* gate_w, const float* gate_b,
                     void* cat_a, void* cat_b, float* lse, float* mask, void* workspace, int64_t workspace_bytes,
                     int nq, int refs, int c, int h, int w_, unsigned flags, void* stream) {
  if (!v_a || !v_b || !w || !gate_w || !cat_a) return COATTN_E_NULL;
  if (nq < 1 || refs < 1) return COATTN_E_SHAPE;
  if (flags & ~(COATTN_FLAG_BF16 | COATTN_FLAG_A_ONLY | COATTN_FLAG_GATED_ONLY)) return COATTN_E_UNSUPPORTED;
  if (refs > 1) flags |= COATTN_FLAG_A_ONLY;        // several references per query frame: frame-A outputs (test.py:301)
  if (!cat_b && !(flags & COATTN_FLAG_A_ONLY)) return COATTN_E_NULL;
  const int n = nq * refs;
  if (int e = check_dims(n, c, h, w_)) return e;
  const Layout ly = make_layout(n, h, w_);
  if (int e = check_workspace(workspace, workspace_bytes, ly)) return e;
  if (int e = check_arch(nullptr)) return e;
  const uintptr_t ptrs = reinterpret_cast<uintptr_t>(v_a) | reinterpret_cast<uintptr_t>(v_b) |
                         reinterpret_cast<uintptr_t>(cat_a) | reinterpret_cast<uintptr_t>(cat_b);
  if (ptrs & 1) return COATTN_E_ALIGN;
  // TMA reads a tensor in place when its base is 16-byte aligned and its rows (L elements) are a multiple of 16 bytes
  const bool in_place = (ly.L % 8 == 0) &&
                        (((reinterpret_cast<uintptr_t>(v_a) | reinterpret_cast<uintptr_t>(v_b)) & 15) == 0);
  const int in16 = in_place ? 2 : 1;
  if (int e = cast_project_core(v_a, v_b, w, reinterpret_cast<unsigned short*>(seg(workspace, ly.off_vv)),
                                reinterpret_cast<unsigned short*>(seg(workspace, ly.off_w16)), n, ly,
                                (flags & COATTN_FLAG_BF16) != 0, uses_fold(flags) ? 2 : 1, static_cast<cudaStream_t>(stream), nq, in16,
                                reinterpret_cast<unsigned*>(seg(workspace, ly.off_status))))
    return e;
  return launch_attend(v_a, v_b, cat_a, cat_b, nullptr, lse, mask, gate_w, gate_b,
                       workspace, workspace_bytes, n, c, h, w_, flags, stream, refs, in16);
}

int coattn_forward(const float* v_a, const float* v_b, const float* w, const float* gate_w, const float* gate_b,
                   float* cat_a, float* cat_b, float* z, float* lse, float* mask, void* workspace,
                   int64_t workspace_bytes, int n, int c, int h, int w_, unsigned flags, void* stream) {
  if (!v_a || !v_b || !w || !gate_w || !cat_a) return COATTN_E_NULL;
  if (!cat_b && !(flags & COATTN_FLAG_A_ONLY)) return COATTN_E_NULL;
  if ((flags & (COATTN_FLAG_A_ONLY | COATTN_FLAG_GATED_ONLY)) && (flags & (COATTN_FLAG_UNFUSED_GATE | COATTN_FLAG_SINGLE_CTA)))
    return COATTN_E_UNSUPPORTED;
  if ((flags & COATTN_FLAG_PLANES_READY) && (flags & (COATTN_FLAG_KMAJOR | COATTN_FLAG_SINGLE_CTA | COATTN_FLAG_UNFUSED_PREP)))
    return COATTN_E_UNSUPPORTED;
  if (int e = check_dims(n, c, h, w_)) return e;
  const Layout ly = make_layout(n, h, w_);
  if (int e = check_workspace(workspace, workspace_bytes, ly)) return e;
  int parts = 1;
  if (int e = split_parts(flags, n, ly, &parts)) return e;
  flags &= ~COATTN_FLAG_SPLIT_KEYS;
  if (parts > 1) {
    if (int e = coattn_stage_prep_project(v_a, v_b, w, workspace, workspace_bytes, n, c, h, w_, flags, stream)) return e;
    return attend_split_merge(v_a, v_b, gate_w, gate_b, cat_a, cat_b, z, lse, mask, workspace, workspace_bytes, n, c, h, w_,
                              flags, stream, 1, parts);
  }
  if (flags & COATTN_FLAG_SINGLE_CTA) flags |= COATTN_FLAG_KMAJOR;   // the cross-check kernel only knows K-major operands
  if (flags & COATTN_FLAG_UNFUSED_PREP) {
    if (int e = coattn_stage_prep(v_a, v_b, w, workspace, workspace_bytes, n, c, h, w_, flags, stream)) return e;
    if (int e = coattn_stage_project(workspace, workspace_bytes, n, c, h, w_, flags, stream)) return e;
  } else {
    if (int e = coattn_stage_prep_project(v_a, v_b, w, workspace, workspace_bytes, n, c, h, w_, flags, stream)) return e;
  }
  if (flags & COATTN_FLAG_UNFUSED_GATE) {
    float* zbuf = z ? z : reinterpret_cast<float*>(seg(workspace, ly.off_z));
    if (int e = coattn_stage_attend(zbuf, lse, workspace, workspace_bytes, n, c, h, w_, flags, stream)) return e;
    return coattn_stage_gate(zbuf, v_a, v_b, gate_w, gate_b, cat_a, cat_b, n, c, h, w_, stream);
  }
  return coattn_stage_attend_gate(v_a, v_b, cat_a, cat_b, z, lse, mask, gate_w, gate_b, workspace, workspace_bytes, n, c,
                                  h, w_, flags, stream);
}


}  // extern "C"

// ---------------------------------------------------------------------------------------------- backward
namespace {
struct BwdLayout {
  Layout fwd;
  int64_t off_g;   // bf16 copy of the plane A16 (operand of dW = dQ A^T, whose other operand dQ is bf16); fp16 forward only
  int64_t off_dza16, off_dzb16, off_delta, off_dta, off_dqt, off_dq16, off_wt, total;
};
// Nothing of size L x L: 16-bit planes [N][C][Lp] and vectors only (the softmax matrices are recomputed tile by tile in
// TMEM, bwd_flash_kernel.cuh).
BwdLayout make_bwd_layout(int n, int h, int w) {
  BwdLayout b{};
  b.fwd = make_layout(n, h, w);
  const int64_t Lp = b.fwd.Lp, L = b.fwd.L;
  const int64_t plane = (int64_t)n * Lp * kC * 2;
  int64_t off = b.fwd.total;
  auto take = [&](int64_t bytes) { const int64_t o = off; off = round_up(off + bytes, kAlign); return o; };
  b.off_g = take(plane);
  b.off_dza16 = take(plane); b.off_dzb16 = take(plane);
  b.off_delta = take((int64_t)2 * n * L * 4);
  b.off_dta = take((int64_t)n * L * 4 + 64);      // d_ta [N][L] + the max |dZ| word
  b.off_dqt = take(plane); b.off_dq16 = take(plane);
  b.off_wt = take((int64_t)kC * kC * 2);
  b.total = off;
  return b;
}

template <int MODE>
int launch_gemm(EncodeTiledFn enc, cudaStream_t st, const void* a, uint64_t a_rows, bool a_bf16, const void* b,
                uint64_t b_rows, bool b_bf16, uint64_t k, int m_tiles, int n_tiles, int batch, GemmParams gp,
                int k_split = 1) {
  CUtensorMap ta, tb;
  if (int e = make_tmap(enc, &ta, a, a_rows, k, 128, a_bf16)) return e;
  if (int e = make_tmap(enc, &tb, b, b_rows, k, 128, b_bf16)) return e;
  gp.num_kb = (int)(k / 64) / k_split;     // callers pick k_split as a divisor of K / 64
  gp.k_split = k_split;
  gp.idesc = (1u << 4) | ((a_bf16 ? 1u : 0u) << 7) | ((b_bf16 ? 1u : 0u) << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  auto kern = gemm_nt_kernel<MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
  if (e != cudaSuccess) return (int)e;
  if ((e = launch_pdl(kern, dim3(m_tiles, n_tiles, batch * k_split), dim3(kNumThreads), kGemmSmemBytes, st, ta, tb, gp)) != cudaSuccess)
    return (int)e;
  return (int)cudaGetLastError();
}

int launch_flash(cudaStream_t st, const FlashMaps& maps, FlashParams& fp, int sms, const Layout& ly, bool fbf16) {
  fp.L = ly.L; fp.Lp = ly.Lp;
  fp.q_pairs = (ly.L + 2 * k2BM - 1) / (2 * k2BM);
  fp.kv_tiles = (ly.L + k2BN - 1) / k2BN;
  const uint32_t n_last = (uint32_t)(((ly.L - (fp.kv_tiles - 1) * k2BN) + 15) & ~15);
  fp.idesc_s = make_idesc_16_major(2 * k2BM, k2BN, fbf16, true, true);
  fp.idesc_s_last = make_idesc_16_major(2 * k2BM, n_last, fbf16, true, true);
  fp.idesc_t = make_idesc_16_major(2 * k2BM, k2BN, fbf16, true, true);
  fp.idesc_t_last = make_idesc_16_major(2 * k2BM, n_last, fbf16, true, true);
  fp.idesc_o = make_idesc_16(2 * k2BM, kC, fbf16);
  // X-producer warps per lane quadrant: 2 (8 warps, 64 tile columns per thread).  COATTN_FLASH_G=4 selects the 16-warp
  // layout of the tuning runs (same arithmetic, identical results): measured SLOWER, 572 vs 540 us for the RGB backward of
  // 8 pairs at 60x60 -- the chain is bound by the MUFU pipe and the hand-offs, not by the warp count (as in attend2).
  static const int g_sel = []() { const char* e = getenv("COATTN_FLASH_G"); return (e && e[0] == '4') ? 4 : 2; }();
  auto kern = g_sel == 2 ? (fbf16 ? bwd_flash_kernel<2, true> : bwd_flash_kernel<2, false>)
                         : (fbf16 ? bwd_flash_kernel<4, true> : bwd_flash_kernel<4, false>);
  const int threads = g_sel == 2 ? FlashCfg<2>::kThreads : FlashCfg<4>::kThreads;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmemBytes);
  if (e != cudaSuccess) return (int)e;
  int clusters = sms / 2;
  if (fp.items[0] + fp.items[1] < clusters) clusters = fp.items[0] + fp.items[1];
  if (clusters < 1) return COATTN_OK;
  if ((e = launch_pdl(kern, dim3(2 * clusters), dim3(threads), kFSmemBytes, st, maps, fp)) != cudaSuccess) return (int)e;
#ifdef COATTN_TRACE_FLASH
  {
    static int calls = 0;
    if (++calls == 3 || calls == 6) {
      long long tb[64 * 16];
      cudaStreamSynchronize(st);
      cudaMemcpyFromSymbol(tb, g_flash_trace, sizeof(tb));
      const long long t0 = tb[0];
      printf("bwd_flash trace (call %d, kinds %d/%d items): cluster 0, second item, first phase\n", calls, fp.items[0], fp.items[1]);
      for (int j = 0; j < 29 && j < fp.kv_tiles; ++j) {
        const long long* r = tb + j * 16;
        printf("tile %2d: top +%7lld | k_full(S j+1) +%5lld | st_free +%5lld | S issued +%5lld | v_full +%5lld | x_full +%5lld | T issued +%5lld"
               " || warp0: S seen +%7lld | in regs +%5lld | P done +%5lld | T seen +%5lld | X stored +%5lld\n", j, r[0] - t0,
               r[13] - r[0], r[14] - r[13], r[1] - r[14], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[8] - t0, r[9] - r[8], r[10] - r[9],
               r[11] - r[10], r[12] - r[11]);
      }
      long long ib[32 * 8];
      cudaMemcpyFromSymbol(ib, g_flash_items, sizeof(ib));
      for (int i = 0; i < 32; ++i) {
        const long long* r = ib + i * 8;
        if (r[0] == 0) break;
        printf("item %2d kind %lld: start +%8lld | sweep A %7lld | B %7lld | C %7lld | wait O %5lld | drain %6lld | total %7lld\n", i, r[7],
               r[0] - ib[0], r[1] - r[0], r[2] ? r[2] - r[1] : 0ll, r[3] ? r[3] - r[2] : 0ll, r[5] - (r[3] ? r[3] : r[2] ? r[2] : r[1]),
               r[6] - r[5], r[6] - r[0]);
      }
    }
  }
#endif
  return (int)cudaGetLastError();
}
}  // namespace

extern "C" {

int64_t coattn_backward_workspace_bytes(int n, int c, int h, int w, int counterpart) {
  (void)counterpart;      // the counterpart-frame gradient needs no extra scratch any more
  if (check_dims(n, c, h, w) != COATTN_OK) return COATTN_E_SHAPE;
  return make_bwd_layout(n, h, w).total;
}

int coattn_backward(const float* v_a, const float* v_b, const float* w, const float* gate_w, const float* z,
                    const float* lse, const float* mask, const float* d_cat_a, const float* d_cat_b, float* d_v_a,
                    float* d_v_b, float* d_w, float* d_gate_w, float* d_gate_b, void* workspace,
                    int64_t workspace_bytes, int n, int c, int h, int w_, unsigned flags, void* stream) {
  if (!v_a || !v_b || !w || !gate_w || !z || !lse || !mask || !d_cat_a || !d_v_a || !d_w || !d_gate_w)
    return COATTN_E_NULL;
  const bool counterpart = d_v_b != nullptr;
  if (int e = check_dims(n, c, h, w_)) return e;
  // d_w is accumulated with 16-byte vector reductions; every other tensor falls back to scalar accesses when unaligned
  if ((reinterpret_cast<uintptr_t>(d_w) & 15) != 0) return COATTN_E_ALIGN;
  const BwdLayout bl = make_bwd_layout(n, h, w_);
  const Layout& ly = bl.fwd;
  if (!workspace) return COATTN_E_NULL;
  if ((reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)) != 0 || workspace_bytes < bl.total) return COATTN_E_WORKSPACE;
  int sms = 148;
  if (int e = check_arch(&sms)) return e;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return COATTN_E_DRIVER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool fbf16 = (flags & COATTN_FLAG_BF16) != 0;   // format of the forward operands (Q16, B16)
  const bool has_b = d_cat_b != nullptr;
  const int L = ly.L, Lp = ly.Lp;
  const int64_t plane_elems = ly.t_pass_elems();

  // Operands of the forward pass are regenerated rather than kept alive between forward and backward: the 16-bit
  // channel-major planes X = [B16, A16, Q16] in the forward's format (S must be recomputed from the operands the forward
  // used, so that exp(S - lse) is the forward's softmax).  The gradient operands of the flash sweeps (dZ_a, dZ_b, X) are in
  // the SAME format -- scaled by one power of two with fp16 (bwd_planes_kernel) -- so the sweeps need no second copy of
  // the features; only dW = dQ A^T, whose dQ is bf16, wants A in bf16.
  // COATTN_FLAG_PLANES_READY: the caller ran the forward of this call on THIS workspace (a backward workspace starts with the
  // forward layout) and nothing has touched it since -- planes B16 and A16 are still there and the feature cast is skipped.
  unsigned short* wt = reinterpret_cast<unsigned short*>(seg(workspace, bl.off_wt));
  unsigned* absmax = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(seg(workspace, bl.off_dta)) + (size_t)n * ly.L);
  {
    unsigned short* w16 = reinterpret_cast<unsigned short*>(seg(workspace, ly.off_w16));
    if (cudaError_t e = launch_pdl(fbf16 ? bwd_init_kernel<true> : bwd_init_kernel<false>, dim3(kC), dim3(kC), 0, st, w, wt, w16, d_w,
                                   d_gate_w, d_gate_b, absmax)) return (int)e;
  }
  if (int e = cast_and_project_mn(v_a, v_b, w, workspace, workspace_bytes, n, c, h, w_,
                                  flags | kInternalNeedQ16 | kInternalW16Ready, stream)) return e;
  unsigned short* xf = reinterpret_cast<unsigned short*>(seg(workspace, ly.off_vv));     // forward format
  unsigned short* b16f = xf;                              // B
  unsigned short* a16f = xf + plane_elems;                // A
  unsigned short* q16f = xf + 2 * plane_elems;            // Q = W A
  unsigned short* a16 = a16f;                             // A, bf16
  if (!fbf16) {
    a16 = reinterpret_cast<unsigned short*>(seg(workspace, bl.off_g));
    CastParams cp;
    cp.va = v_a; cp.vb = v_b;
    cp.x = a16 - plane_elems;      // the kernel writes plane 1 of x
    cp.N = n; cp.L = L; cp.Lp = Lp; cp.Na = n; cp.status = nullptr; cp.first_plane = 1; cp.w = nullptr; cp.w16 = nullptr;
    const bool vec = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(v_a) & 15) == 0);
    if (cudaError_t e = launch_pdl(vec ? cast_kernel<true, 4> : cast_kernel<true, 1>, dim3(n * kC, 1), dim3(256), 0, st, cp)) return (int)e;
  }
  unsigned short* dza16 = reinterpret_cast<unsigned short*>(seg(workspace, bl.off_dza16));
  unsigned short* dzb16 = reinterpret_cast<unsigned short*>(seg(workspace, bl.off_dzb16));
  float* delta = reinterpret_cast<float*>(seg(workspace, bl.off_delta));
  float* d_ta = reinterpret_cast<float*>(seg(workspace, bl.off_dta));
  unsigned short* dqt = reinterpret_cast<unsigned short*>(seg(workspace, bl.off_dqt));
  unsigned short* dq16 = reinterpret_cast<unsigned short*>(seg(workspace, bl.off_dq16));

  cudaError_t ce;

  BwdPrepParams bp;
  bp.d_cat_a = d_cat_a; bp.d_cat_b = d_cat_b; bp.z = z; bp.mask = mask; bp.gate_w = gate_w;
  bp.d_vb = d_v_b;
  bp.dza16 = dza16; bp.dzb16 = dzb16; bp.delta = delta; bp.d_ta = d_ta; bp.absmax = absmax;
  bp.d_gate_w = d_gate_w; bp.d_gate_b = d_gate_b; bp.d_va = d_v_a; bp.N = n; bp.L = L; bp.Lp = Lp;
  bp.cat_ch = (flags & COATTN_FLAG_GATED_ONLY) ? kC : 2 * kC;
  if ((ce = launch_pdl(bwd_stats_kernel, dim3(Lp / kBwdPrepPos, n), dim3(kBwdPrepThreads), 0, st, bp)) != cudaSuccess) return (int)ce;
  if ((ce = launch_pdl(fbf16 ? bwd_planes_kernel<true> : bwd_planes_kernel<false>, dim3(Lp / kBwdPrepPos, n), dim3(kBwdPrepThreads), 0, st,
                       bp)) != cudaSuccess) return (int)ce;
  if ((ce = cudaGetLastError()) != cudaSuccess) return (int)ce;

  const uint64_t rowsL = (uint64_t)n * Lp, rowsC = (uint64_t)n * kC;
  const int lt = Lp / 128;
  // ---- the L x L part: flash-style sweeps, nothing of that size is written (bwd_flash_kernel.cuh)
  FlashMaps maps;
  enum { mQf = 0, mBf = 1, mDza = 2, mBg = 3, mAg = 4, mDzb = 5, vBg = 6, vDzb = 7, vQg = 8, vDza = 9 };
  {
    // every operand of the sweeps is in the forward's format (the "g" maps used to be bf16 copies)
    struct { int idx; const void* base; uint32_t box; } defs[kFMaps] = {
        {mQf, q16f, 256}, {mBf, b16f, 256}, {mDza, dza16, 256}, {mBg, b16f, 256}, {mAg, a16f, 256},
        {mDzb, dzb16, 256}, {vBg, b16f, 128}, {vDzb, dzb16, 128}, {vQg, q16f, 128}, {vDza, dza16, 128}};
    for (const auto& d : defs)
      if (int e = make_tmap(enc, &maps.m[d.idx], d.base, rowsC, Lp, d.box, fbf16)) return e;
  }
  const float* lse_a = lse;
  const float* lse_b = lse + (size_t)n * L;
  const float* del_a = delta;
  const float* del_b = delta + (size_t)n * L;
  const int items = n * ((L + 2 * k2BM - 1) / (2 * k2BM));
  {
    // frame-A side: dQ (phase A [+ phase B]) and dA += P_b dZ_b^T
    FlashParams fp{};
    fp.N = n;
    fp.absmax = fbf16 ? nullptr : absmax;
    // S and dP_a both multiply B_J.  COATTN_FLASH_SHARED=1 loads that tile ONCE per column tile (c1 == c2: the two ring
    // slots then alternate, double buffered); measured 1 % SLOWER than two loads of the same tile (560 vs 556 us, RGB
    // backward of 8 pairs at 60x60, A/B on one box): the sweep is bound by the T -> X -> PV -> T chain, not by the loads.
    static const bool share = []() { const char* e = getenv("COATTN_FLASH_SHARED"); return e && e[0] == '1'; }();
    fp.ph[0] = FlashPhase{mQf, mBf, mDza, share ? mBf : mBg, vBg, 0, lse_a, del_a};
    fp.ph[1] = FlashPhase{mQf, mBf, mAg, mDzb, vBg, 1, lse_b, del_b};
    fp.ph[2] = FlashPhase{mQf, mBf, -1, -1, vDzb, 1, lse_b, nullptr};
    fp.kind[0] = FlashKind{0, has_b ? 2 : 1, 0, dqt, dq16, nullptr};
    fp.kind[1] = FlashKind{2, 1, 1, nullptr, nullptr, d_v_a};
    fp.items[0] = items;
    fp.items[1] = has_b ? items : 0;
    fp.ratio = 3;      // two phases of three products against one phase of two
    if (int e = launch_flash(st, maps, fp, sms, ly, fbf16)) return e;
  }
  if (counterpart) {
    // frame-B side (no_grad_for_counterpart=False): the same sweeps with the roles of the frames swapped, all into d_v_b
    //   dB[:, j] += sum_i dS[i, j] Q[:, i] + sum_i P_a[i, j] dZ_a[:, i]
    FlashParams fp{};
    fp.N = n;
    fp.absmax = fbf16 ? nullptr : absmax;
    fp.ph[0] = FlashPhase{mBf, mQf, mBg, mDza, vQg, 1, lse_a, del_a};        // P_a (dP_a - delta_a): vectors follow the columns (i)
    fp.ph[1] = FlashPhase{mBf, mQf, -1, -1, vDza, 1, lse_a, nullptr};         // P_a dZ_a
    fp.ph[2] = FlashPhase{mBf, mQf, mDzb, mAg, vQg, 0, lse_b, del_b};        // P_b (dP_b - delta_b): vectors follow the rows (j)
    fp.kind[0] = FlashKind{0, has_b ? 3 : 2, 1, nullptr, nullptr, d_v_b};
    fp.kind[1] = FlashKind{0, 0, 1, nullptr, nullptr, nullptr};
    fp.items[0] = items;
    fp.items[1] = 0;
    fp.ratio = 1;
    if (int e = launch_flash(st, maps, fp, sms, ly, fbf16)) return e;
  }
  GemmParams gp{};
  // dA[c][i] += sum_d dQt[i][d] W[d][c]
  gp.out0 = d_v_a; gp.ld0 = L; gp.rows0 = kC; gp.out1 = nullptr; gp.m_valid = L;
  gp.a_rows_per_batch = Lp; gp.b_rows_per_batch = 0;
  if (int e = launch_gemm<kGemmAddF32T>(enc, st, dqt, rowsL, true, wt, kC, true, kC, lt, kC / 128, n, gp)) return e;
  // dW[d][c] += sum_n sum_i dQ16[n][d][i] A16[n][c][i]
  gp.out0 = d_w; gp.ld0 = kC; gp.rows0 = 0; gp.m_valid = kC;
  gp.a_rows_per_batch = kC; gp.b_rows_per_batch = kC;
  // only 4 output tiles per sample: split the position range over several CTAs (the atomics already reduce over n)
  int dw_split = 1;
  for (int cand = 2; cand <= 8; ++cand)
    if ((Lp / 64) % cand == 0 && 4 * n * cand <= 2 * sms) dw_split = cand;
  if (int e = launch_gemm<kGemmAtomicF32>(enc, st, dq16, rowsC, true, a16, rowsC, true, Lp, kC / 128, kC / 128, n, gp, dw_split))
    return e;
  return COATTN_OK;
}

}  // extern "C"
