// Data-parallel optimizer step over NVLink peer memory: gradient reduction, global-norm clip, AdamW and the parameter
// broadcast of a replicated model as TWO kernels that read / write the other GPUs' buffers directly, in place of
// "NCCL all-reduce of the whole gradient, then AdamW over the whole model on every rank".
//
// What the reference's notebooks get from accelerate / DDP (Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1 main():
// DDP gradient all-reduce, clip_grad_norm_(1.0), AdamW) is, per step and rank, 2 (N-1)/N x |grad| of NVLink traffic plus
// 28 bytes of HBM traffic per parameter for an optimizer that every rank repeats identically. Here rank r owns the
// contiguous shard [lo, hi) of the flat parameter buffer:
//   vy_dp_reduce_shard   g[i] = sum over ranks of grad_rank[i] for i in the shard — peer loads through NVLink, fixed rank
//                        order, fp32 accumulation — kept in fp32; the shard's sum of squares is published to every rank;
//   vy_dp_adamw_shard    clip coefficient from the sum of the N published partial norms (fixed order: identical on every
//                        rank), AdamW on the shard (fp32 master weights and moments exist ONLY for the shard: 1/N of the
//                        optimizer state and of its HBM traffic), and the new bf16/fp32 parameters are stored into every
//                        rank's parameter buffer through NVLink.
// vy_dp_barrier is the device-side rendezvous between the phases (a flag per peer in symmetric memory, generation
// counted), so the whole step stays inside one captured CUDA graph. Every rank ends the step with bit-identical parameters.
#include "vy_common.cuh"
#include "vy_ptx.cuh"

namespace vy {

constexpr int DP_MAX = 8;

struct DpPeersDev {
  int world, rank;
  const void* grads[DP_MAX];
  void* params[DP_MAX];
  unsigned int* flags[DP_MAX];
  float* scalars[DP_MAX];
  unsigned int* epoch;
  int* error_flag;
};

static int load_group(const VyDpGroup* g, DpPeersDev* d, const char* who) {
  VY_CHECK_ARG(g != nullptr, "%s: null group", who);
  VY_CHECK_ARG(g->world >= 1 && g->world <= DP_MAX && g->rank >= 0 && g->rank < g->world, "%s: world %d / rank %d outside [1, %d]", who,
               g->world, g->rank, DP_MAX);
  VY_CHECK_ARG(g->flags && g->epoch && g->error_flag, "%s: null barrier state", who);
  memset(d, 0, sizeof(*d));
  d->world = g->world;
  d->rank = g->rank;
  for (int r = 0; r < g->world; ++r) {
    d->grads[r] = g->grads ? g->grads[r] : nullptr;
    d->params[r] = g->params ? g->params[r] : nullptr;
    d->flags[r] = g->flags[r];
    d->scalars[r] = g->scalars ? g->scalars[r] : nullptr;
    VY_CHECK_ARG(d->flags[r] != nullptr, "%s: null flags pointer of rank %d", who, r);
  }
  d->epoch = g->epoch;
  d->error_flag = g->error_flag;
  return VY_OK;
}

__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One CTA. Thread t < world raises flag[rank] on peer t to this rank's new generation, then waits until peer t has raised
// flag[t] here. Everything the calling stream did before is visible to the peers after they leave the barrier.
__global__ void __launch_bounds__(32)
dp_barrier_kernel(const DpPeersDev g) {
  __shared__ unsigned int s_gen;
  if (threadIdx.x == 0) {
    s_gen = *g.epoch + 1u;
    *g.epoch = s_gen;
    __threadfence_system();
  }
  __syncwarp();
  const unsigned int gen = s_gen;
  const int t = threadIdx.x;
  if (t < g.world) {
    st_release_sys_u32(&g.flags[t][g.rank], gen);
    unsigned int spins = 0;
    // (generations only grow; signed distance tolerates the 2^32 wrap)
    while (static_cast<int>(ld_acquire_sys_u32(&g.flags[g.rank][t]) - gen) < 0) {
      if (++spins > (1u << 26)) {  // seconds: a peer never arrived — flag it instead of hanging the GPU
        atomicExch(g.error_flag, 1);
        break;
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}

// deterministic sum of squares of the shard (block partials, last block sums them in index order)
__device__ float dp_sq_partials[1024];
__device__ unsigned int dp_sq_ticket = 0;

struct DpReduceDev {
  long long lo, hi;  // shard, multiples of 8
  int dtype;
  float* gshard;     // [hi - lo] fp32
  float* sq_local;   // device scalar: this shard's sum of squares
};

__global__ void __launch_bounds__(256)
dp_reduce_kernel(const DpPeersDev g, const DpReduceDev p) {
  __shared__ float s_red[8];
  __shared__ bool s_last;
  float a = 0.f;
  const long long nvec = (p.hi - p.lo) >> 3;
  for (long long vi = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; vi < nvec;
       vi += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = p.lo + vi * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // all peers' loads first (independent NVLink round trips), then the sum in rank order
    float v[DP_MAX][8];
#pragma unroll
    for (int r = 0; r < DP_MAX; ++r)
      if (r < g.world) ld8_as_float(g.grads[r], p.dtype, i, v[r]);
#pragma unroll
    for (int r = 0; r < DP_MAX; ++r)
      if (r < g.world) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[r][j];
      }
    st8_from_float(p.gshard, VY_F32, vi * 8, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) a += acc[j] * acc[j];
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w];
    dp_sq_partials[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(&dp_sq_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += 32) t += *const_cast<volatile float*>(&dp_sq_partials[i]);
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      *p.sq_local = t;
      dp_sq_ticket = 0;
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    // publish this shard's sum of squares in slot [rank] of every rank (read after the next barrier)
    if (threadIdx.x < g.world) {
      g.scalars[threadIdx.x][g.rank] = t;
      __threadfence_system();
    }
  }
}

struct DpAdamWDev {
  long long lo, hi;
  int p_dt;
  const float* gshard;
  float* m;
  float* v;
  float* master;  // fp32 copy of the shard, or null when the parameters are fp32 themselves
  float lr, beta1, beta2, eps, wd, max_norm;
  const int* step_ptr;
  int step;
};

__global__ void __launch_bounds__(256)
dp_adamw_kernel(const DpPeersDev g, const DpAdamWDev p) {
  // global gradient norm of the MEAN gradient: sum of the ranks' published shard norms, in rank order
  float sq = 0.f;
  for (int r = 0; r < g.world; ++r) sq += __ldcg(&g.scalars[g.rank][r]);
  const float inv_world = 1.f / static_cast<float>(g.world);
  float clip = inv_world;
  if (p.max_norm > 0.f) {
    const float norm = sqrtf(sq) * inv_world;
    clip *= fminf(1.f, p.max_norm / (norm + 1e-6f));
  }
  const float t = p.step_ptr ? static_cast<float>(*p.step_ptr) : static_cast<float>(p.step);
  const float inv_bc1 = 1.f / (1.f - powf(p.beta1, t)), inv_bc2 = 1.f / (1.f - powf(p.beta2, t));
  const float decay = 1.f - p.lr * p.wd;
  const long long nvec = (p.hi - p.lo) >> 3;
  for (long long vi = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; vi < nvec;
       vi += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long li = vi * 8, i = p.lo + li;
    float gi[8], w[8], mi[8], vv[8];
    ld8_as_float(p.gshard, VY_F32, li, gi);
    if (p.master) ld8_as_float(p.master, VY_F32, li, w);
    else ld8_as_float(g.params[g.rank], p.p_dt, i, w);
    ld8_as_float(p.m, VY_F32, li, mi);
    ld8_as_float(p.v, VY_F32, li, vv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gj = gi[j] * clip;
      mi[j] = p.beta1 * mi[j] + (1.f - p.beta1) * gj;
      vv[j] = p.beta2 * vv[j] + (1.f - p.beta2) * gj * gj;
      w[j] = w[j] * decay - p.lr * (mi[j] * inv_bc1) / (sqrtf(vv[j] * inv_bc2) + p.eps);
    }
    st8_from_float(p.m, VY_F32, li, mi);
    st8_from_float(p.v, VY_F32, li, vv);
    if (p.master) st8_from_float(p.master, VY_F32, li, w);
    // the updated parameters go to every rank's buffer (own copy included): NVLink stores
#pragma unroll
    for (int r = 0; r < DP_MAX; ++r)
      if (r < g.world) st8_from_float(g.params[r], p.p_dt, i, w);
  }
}

static int dp_grid(long long nvec) {
  long long blocks = (nvec + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 4;
  if (blocks > cap) blocks = cap;
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace vy

using namespace vy;

extern "C" int vy_dp_barrier(const VyDpGroup* g, void* stream) {
  if (!vy_device_ok()) {
    set_error("vy_dp_barrier: no sm_100 device (there is no CPU fallback)");
    return VY_ERR_NO_DEVICE;
  }
  DpPeersDev d;
  int rc = load_group(g, &d, "vy_dp_barrier");
  if (rc != VY_OK) return rc;
  VY_CUDA_OK(launch_kernel(dp_barrier_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream), d));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_dp_reduce_shard(const VyDpGroup* g, const VyDpReduce* p) {
  if (!vy_device_ok()) {
    set_error("vy_dp_reduce_shard: no sm_100 device (there is no CPU fallback)");
    return VY_ERR_NO_DEVICE;
  }
  VY_CHECK_ARG(p != nullptr, "vy_dp_reduce_shard: null params");
  DpPeersDev d;
  int rc = load_group(g, &d, "vy_dp_reduce_shard");
  if (rc != VY_OK) return rc;
  VY_CHECK_ARG(p->lo >= 0 && p->hi > p->lo && (p->lo & 7) == 0 && (p->hi & 7) == 0, "vy_dp_reduce_shard: shard [%lld, %lld) must be non-empty, multiples of 8",
               (long long)p->lo, (long long)p->hi);
  VY_CHECK_ARG(dtype_ok(p->dtype) && p->gshard && p->sq_local, "vy_dp_reduce_shard: null pointer / bad dtype");
  for (int r = 0; r < d.world; ++r) VY_CHECK_ARG(d.grads[r] && d.scalars[r], "vy_dp_reduce_shard: rank %d: null gradient / scalar pointer", r);
  DpReduceDev q;
  q.lo = p->lo; q.hi = p->hi; q.dtype = p->dtype; q.gshard = p->gshard; q.sq_local = p->sq_local;
  VY_CUDA_OK(launch_kernel(dp_reduce_kernel, dim3(dp_grid((p->hi - p->lo) >> 3)), dim3(256), 0, static_cast<cudaStream_t>(p->stream), d, q));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}

extern "C" int vy_dp_adamw_shard(const VyDpGroup* g, const VyDpAdamW* p) {
  if (!vy_device_ok()) {
    set_error("vy_dp_adamw_shard: no sm_100 device (there is no CPU fallback)");
    return VY_ERR_NO_DEVICE;
  }
  VY_CHECK_ARG(p != nullptr, "vy_dp_adamw_shard: null params");
  DpPeersDev d;
  int rc = load_group(g, &d, "vy_dp_adamw_shard");
  if (rc != VY_OK) return rc;
  VY_CHECK_ARG(p->lo >= 0 && p->hi > p->lo && (p->lo & 7) == 0 && (p->hi & 7) == 0, "vy_dp_adamw_shard: shard must be non-empty, multiples of 8");
  VY_CHECK_ARG(dtype_ok(p->param_dtype) && p->gshard && p->exp_avg && p->exp_avg_sq, "vy_dp_adamw_shard: null pointer / bad dtype");
  VY_CHECK_ARG(p->step >= 1 || p->step_ptr, "vy_dp_adamw_shard: step must be >= 1 (or pass step_ptr)");
  for (int r = 0; r < d.world; ++r) VY_CHECK_ARG(d.params[r] && d.scalars[r], "vy_dp_adamw_shard: rank %d: null parameter / scalar pointer", r);
  DpAdamWDev q;
  q.lo = p->lo; q.hi = p->hi; q.p_dt = p->param_dtype; q.gshard = p->gshard; q.m = p->exp_avg; q.v = p->exp_avg_sq; q.master = p->master;
  q.lr = p->lr; q.beta1 = p->beta1; q.beta2 = p->beta2; q.eps = p->eps; q.wd = p->weight_decay; q.max_norm = p->max_grad_norm;
  q.step_ptr = p->step_ptr; q.step = p->step;
  VY_CUDA_OK(launch_kernel(dp_adamw_kernel, dim3(dp_grid((p->hi - p->lo) >> 3)), dim3(256), 0, static_cast<cudaStream_t>(p->stream), d, q));
  VY_LAUNCH_OK();
  count_launch();
  return VY_OK;
}
