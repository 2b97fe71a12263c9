// Two boundary entry points SURVEY §8 b4 names beyond the text plan:
//   * lagvae_comm_* / lagvae_allreduce_bucket — the ONE data-path collective of the sharded inner step: a sum
//     all-reduce of the flat fp32 gradient bucket over NCCL, the communicator owned by the library (no torch types).  NCCL is
//     bound at run time with dlopen (the path of torch's bundled libnccl.so.2 is passed in by the host side), so liblagvae.so
//     itself has no link-time dependency on it and loads on boxes without NCCL.
//   * lagvae_adam_table_* / lagvae_clip_adam_step — clip_grad_norm_(all params) + torch.optim.Adam step on the first n_update
//     tensors (image.py:312-314; optim.Adam(lr=1e-3) image.py:267) over an arbitrary number of parameter tensors (the image
//     model has 248): a device-resident segment table built once, then three launches per step (capturable in a CUDA graph:
//     the step count for the bias correction lives on the device).
#include <dlfcn.h>

#include <cmath>
#include <new>
#include <vector>

#include "kernels.cuh"

using namespace lagvae;

// ---------------------------------------------------------------------------------------------------------------------
// NCCL through dlopen
// ---------------------------------------------------------------------------------------------------------------------
namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
typedef int (*pfn_getuid)(nccl_uid*);
typedef int (*pfn_initrank)(nccl_comm*, int, nccl_uid, int);
typedef int (*pfn_allreduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*pfn_destroy)(nccl_comm);
typedef const char* (*pfn_errstr)(int);
struct NcclApi {
  void* handle = nullptr;
  pfn_getuid get_uid = nullptr;
  pfn_initrank init_rank = nullptr;
  pfn_allreduce all_reduce = nullptr;
  pfn_destroy destroy = nullptr;
  pfn_errstr errstr = nullptr;
} g_nccl;
constexpr int NCCL_FLOAT32 = 7, NCCL_SUM = 0;   // ncclDataType_t / ncclRedOp_t values of nccl.h (stable across 2.x)
}  // namespace

struct lagvae_comm {
  nccl_comm comm;
  int rank, world;
};

extern "C" {

int lagvae_comm_load(const char* libnccl_path) {
  if (g_nccl.handle) return LAGVAE_OK;
  LV_CHECK_ARG(libnccl_path != nullptr, "comm_load: null path");
  void* h = dlopen(libnccl_path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("comm_load: dlopen(%s) failed: %s", libnccl_path, dlerror());
    return LAGVAE_E_ARG;
  }
  g_nccl.get_uid = (pfn_getuid)dlsym(h, "ncclGetUniqueId");
  g_nccl.init_rank = (pfn_initrank)dlsym(h, "ncclCommInitRank");
  g_nccl.all_reduce = (pfn_allreduce)dlsym(h, "ncclAllReduce");
  g_nccl.destroy = (pfn_destroy)dlsym(h, "ncclCommDestroy");
  g_nccl.errstr = (pfn_errstr)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.all_reduce || !g_nccl.destroy) {
    set_error("comm_load: %s does not export the NCCL entry points", libnccl_path);
    dlclose(h);
    return LAGVAE_E_ARG;
  }
  g_nccl.handle = h;
  return LAGVAE_OK;
}

int lagvae_comm_unique_id(void* out_128_bytes) {
  LV_CHECK_ARG(g_nccl.handle, "comm_unique_id: call lagvae_comm_load first");
  LV_CHECK_ARG(out_128_bytes != nullptr, "comm_unique_id: null output");
  const int r = g_nccl.get_uid((nccl_uid*)out_128_bytes);
  if (r != 0) {
    set_error("ncclGetUniqueId: %s", g_nccl.errstr ? g_nccl.errstr(r) : "error");
    return LAGVAE_E_CUDA;
  }
  return LAGVAE_OK;
}

int lagvae_comm_init(const void* uid_128_bytes, int rank, int world, lagvae_comm** out) {
  LV_CHECK_ARG(g_nccl.handle, "comm_init: call lagvae_comm_load first");
  LV_CHECK_ARG(uid_128_bytes && out && world >= 1 && rank >= 0 && rank < world, "comm_init: bad argument");
  lagvae_comm* c = new (std::nothrow) lagvae_comm{};
  LV_CHECK_ARG(c != nullptr, "comm_init: host allocation failed");
  nccl_uid uid;
  memcpy(&uid, uid_128_bytes, sizeof(uid));
  const int r = g_nccl.init_rank(&c->comm, world, uid, rank);
  if (r != 0) {
    set_error("ncclCommInitRank: %s", g_nccl.errstr ? g_nccl.errstr(r) : "error");
    delete c;
    return LAGVAE_E_CUDA;
  }
  c->rank = rank;
  c->world = world;
  *out = c;
  return LAGVAE_OK;
}

int lagvae_allreduce_bucket(lagvae_comm* comm, float* bucket, int64_t count, void* stream) {
  LV_CHECK_ARG(comm && bucket && count >= 0, "allreduce_bucket: bad argument");
  if (count == 0 || comm->world == 1) return LAGVAE_OK;
  const int r = g_nccl.all_reduce(bucket, bucket, (size_t)count, NCCL_FLOAT32, NCCL_SUM, comm->comm, (cudaStream_t)stream);
  if (r != 0) {
    set_error("ncclAllReduce: %s", g_nccl.errstr ? g_nccl.errstr(r) : "error");
    return LAGVAE_E_CUDA;
  }
  return LAGVAE_OK;
}

void lagvae_comm_destroy(lagvae_comm* comm) {
  if (!comm) return;
  if (g_nccl.destroy) g_nccl.destroy(comm->comm);
  delete comm;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// clip + Adam over a segment table
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int CHUNK = 4096;          // elements per work item
struct Seg {
  float *p, *g, *m, *v;
  int64_t n;
};
struct Work {
  int seg;
  int64_t off;
};
}  // namespace

struct lagvae_adam_table {
  Seg* d_seg;
  Work* d_work;
  double* d_partial;
  float* d_coef;       // [0] clip coefficient, [1] pre-clip norm
  int* d_step;
  int n_seg, n_update, n_work, n_work_update;
};

namespace {
__global__ void __launch_bounds__(256) k_tab_sumsq(const Seg* __restrict__ seg, const Work* __restrict__ work, int n_work,
                                                   double* __restrict__ partial) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
    const Seg s = seg[work[w].seg];
    const int64_t o = work[w].off;
    const int64_t e = min(o + (int64_t)CHUNK, s.n);
    for (int64_t i = o + threadIdx.x; i < e; i += blockDim.x) {
      const float x = s.g[i];
      acc += x * x;
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = (double)acc;
}
__global__ void __launch_bounds__(256) k_tab_finish(const double* __restrict__ partial, int n, float max_norm, float* __restrict__ coef,
                                                    float* __restrict__ out_norm, int* __restrict__ step) {
  __shared__ double red[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += partial[i];   // fixed order: deterministic
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float nrm = (float)sqrt(red[0]);
    const float c = max_norm / (nrm + 1e-6f);                           // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max = 1)
    coef[0] = c < 1.f ? c : 1.f;
    coef[1] = nrm;
    if (out_norm) *out_norm = nrm;
    *step += 1;                                                         // Adam's state['step'] += 1
  }
}
// torch.optim.Adam (amsgrad = False, weight_decay = 0, maximize = False), single-tensor formulation:
//   m = b1 m + (1 - b1) g ; v = b2 v + (1 - b2) g^2 ; p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void __launch_bounds__(256) k_tab_clip_adam(const Seg* __restrict__ seg, const Work* __restrict__ work, int n_work, int n_update,
                                                       const float* __restrict__ coef_p, const int* __restrict__ step_p, float lr,
                                                       float b1, float b2, float eps, int scale_all) {
  const float coef = coef_p[0];
  const int t = *step_p;
  const float bc1 = 1.f - powf(b1, (float)t), bc2 = 1.f - powf(b2, (float)t);
  const float step_size = lr / bc1, inv_bc2_sqrt = rsqrtf(bc2);
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
    const int si = work[w].seg;
    const bool upd = si < n_update;
    if (!upd && !scale_all) continue;
    const Seg s = seg[si];
    const int64_t o = work[w].off;
    const int64_t e = min(o + (int64_t)CHUNK, s.n);
    for (int64_t i = o + threadIdx.x; i < e; i += blockDim.x) {
      const float g = s.g[i] * coef;                                    // clip_grad_norm_ multiplies in place (also when coef == 1)
      s.g[i] = g;
      if (upd) {
        const float m = b1 * s.m[i] + (1.f - b1) * g;
        const float v = b2 * s.v[i] + (1.f - b2) * g * g;
        s.m[i] = m;
        s.v[i] = v;
        s.p[i] -= step_size * (m / (sqrtf(v) * inv_bc2_sqrt + eps));
      }
    }
  }
}
}  // namespace

extern "C" {

size_t lagvae_adam_table_bytes(const int64_t* h_counts, int n_seg) {
  if (!h_counts || n_seg <= 0) return 0;
  int64_t nw = 0;
  for (int i = 0; i < n_seg; ++i) nw += (h_counts[i] + CHUNK - 1) / CHUNK;
  return (size_t)round_up((int64_t)(n_seg * sizeof(Seg) + nw * sizeof(Work) + 148 * 8 * sizeof(double) + 64), 256) + 256;
}

int lagvae_adam_table_create(float* const* h_params, float* const* h_grads, float* const* h_exp_avg, float* const* h_exp_avg_sq,
                             const int64_t* h_counts, int n_seg, int n_update, int initial_step, void* device_mem,
                             size_t device_bytes, void* stream, lagvae_adam_table** out) {
  LV_CHECK_ARG(h_grads && h_counts && out && device_mem && n_seg > 0 && n_update >= 0 && n_update <= n_seg, "adam_table_create: bad argument");
  LV_CHECK_ARG(n_update == 0 || (h_params && h_exp_avg && h_exp_avg_sq), "adam_table_create: null parameter / state arrays");
  LV_CHECK_ARG(device_bytes >= lagvae_adam_table_bytes(h_counts, n_seg), "adam_table_create: device memory too small");
  LV_CHECK_ARG(((uintptr_t)device_mem & 255) == 0, "adam_table_create: device memory must be 256-B aligned");
  std::vector<Seg> segs(n_seg);
  std::vector<Work> work;
  int n_work_update = 0;
  for (int i = 0; i < n_seg; ++i) {
    segs[i] = Seg{i < n_update ? h_params[i] : nullptr, h_grads[i], i < n_update ? h_exp_avg[i] : nullptr,
                  i < n_update ? h_exp_avg_sq[i] : nullptr, h_counts[i]};
    for (int64_t o = 0; o < h_counts[i]; o += CHUNK) work.push_back(Work{i, o});
    if (i == n_update - 1) n_work_update = (int)work.size();
  }
  lagvae_adam_table* t = new (std::nothrow) lagvae_adam_table{};
  LV_CHECK_ARG(t != nullptr, "adam_table_create: host allocation failed");
  char* p = (char*)device_mem;
  t->d_seg = (Seg*)p;
  p += round_up((int64_t)(n_seg * sizeof(Seg)), 16);
  t->d_work = (Work*)p;
  p += round_up((int64_t)(work.size() * sizeof(Work)), 16);
  t->d_partial = (double*)p;
  p += 148 * 8 * sizeof(double);
  t->d_coef = (float*)p;
  t->d_step = (int*)(p + 16);
  t->n_seg = n_seg;
  t->n_update = n_update;
  t->n_work = (int)work.size();
  t->n_work_update = n_work_update;
  cudaStream_t st = (cudaStream_t)stream;
  // synchronous uploads of small host tables (one-time set-up, outside any graph capture)
  LV_CUDA(cudaMemcpyAsync(t->d_seg, segs.data(), n_seg * sizeof(Seg), cudaMemcpyHostToDevice, st));
  LV_CUDA(cudaMemcpyAsync(t->d_work, work.data(), work.size() * sizeof(Work), cudaMemcpyHostToDevice, st));
  LV_CUDA(cudaMemcpyAsync(t->d_step, &initial_step, sizeof(int), cudaMemcpyHostToDevice, st));
  LV_CUDA(cudaStreamSynchronize(st));
  *out = t;
  return LAGVAE_OK;
}

void lagvae_adam_table_destroy(lagvae_adam_table* t) { delete t; }

int lagvae_clip_adam_step(lagvae_adam_table* t, float max_norm, float lr, float beta1, float beta2, float eps,
                          int scale_all_grads, float* out_norm, void* stream) {
  LV_CHECK_ARG(t != nullptr, "clip_adam_step: null table");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = 148 * 8;
  k_tab_sumsq<<<nblk, 256, 0, st>>>(t->d_seg, t->d_work, t->n_work, t->d_partial);
  LV_LAUNCH_CHECK();
  k_tab_finish<<<1, 256, 0, st>>>(t->d_partial, nblk, max_norm, t->d_coef, out_norm, t->d_step);
  LV_LAUNCH_CHECK();
  k_tab_clip_adam<<<148 * 4, 256, 0, st>>>(t->d_seg, t->d_work, scale_all_grads ? t->n_work : t->n_work_update, t->n_update, t->d_coef,
                                           t->d_step, lr, beta1, beta2, eps, scale_all_grads);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

}  // extern "C"
