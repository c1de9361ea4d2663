// C-ABI glue: error reporting, device check, and the fused per-chunk step of variant R.
#include <stdarg.h>

#include <time.h>

#include "common.cuh"

namespace ltm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

__global__ void arch_probe_kernel(int* out) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_SPECIFIC__) || (__CUDA_ARCH__ >= 1000)
  *out = __CUDA_ARCH__;
#else
  *out = -1;
#endif
}

}  // namespace ltm

extern "C" int ltm_version(void) { return 100; }   // 0.1.0

extern "C" const char* ltm_last_error(void) { return ltm::g_err; }

extern "C" int ltm_device_check(void) {
  using namespace ltm;
  int dev = 0;
  LTM_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  LTM_CUDA(cudaGetDeviceProperties(&prop, dev));
  LTM_REQUIRE(prop.major == 10, "device %s is sm_%d%d; libinfltm carries sm_100a code only", prop.name, prop.major,
              prop.minor);
  cudaFuncAttributes fa;
  LTM_CUDA(cudaFuncGetAttributes(&fa, arch_probe_kernel));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// One chunk of variant R for Bv videos:  pool -> (re-sample) -> consolidate -> project K,V ->
// continuous attention (+ next call's sticky histogram).  long_term_attention_gibbs.py:288-346.
// ------------------------------------------------------------------------------------------------
extern "C" int ltm_rect_step(const ltm_rect_step_args* a, const float* k, const float* q, const double* u,
                             const uint8_t* new_doc, float* ctx, void* stream) {
  using namespace ltm;
  LTM_REQUIRE(a && q && ctx, "rect_step: null pointer");
  const int D = a->H * a->d;
  const int qtiles = (a->Q + 31) / 32;
  cudaStream_t st_ = (cudaStream_t)stream;
#define LTM_PROF(i) do { if (a->prof_events[i]) cudaEventRecord((cudaEvent_t)a->prof_events[i], st_); } while (0)
  int rc = 0;
  const float* B_past = a->B_past;
  // binned: a->xpart carries per-bin sums of the pooled frames (ltm_pool_bins) and the update tables list bins
  const bool binned = a->binned != 0;
  if (binned) {
    LTM_REQUIRE(B_past != nullptr && new_doc == nullptr, "rect_step: bin sums only exist for update calls of all videos");
    LTM_REQUIRE(a->xb_rows > 0 && a->fbin_ptr && a->seg_ptr1b && a->seg_mem1b, "rect_step: binned tables missing");
  }
  if (k != nullptr) {          // k == NULL: a->xpart was already filled (frame pooling is stateless and may be
    LTM_PROF(0);               // issued ahead of time on another stream, see BatchedRectLTM.prefetch)
    rc = binned ? ltm_pool_bins(k, a->xpart, a->fbin_ptr, a->Bv, a->L, a->T, a->e, a->xb_rows, stream)
                : ltm_pool_mean(k, a->xpart, a->Bv, a->L, a->T, a->e, a->splits, stream);
    if (rc) return rc;
    LTM_PROF(1);
  }
  const int xL = binned ? a->xb_rows : a->L, xsplits = binned ? 1 : a->splits;
  const int32_t* seg_ptr1 = binned ? a->seg_ptr1b : a->seg_ptr1;
  const int32_t* seg_mem1 = binned ? a->seg_mem1b : a->seg_mem1;
  const int32_t* idx = a->idx_uniform;     // non-sticky: fixed table shared by all videos (index stride 0)
  long long idx_stride = 0;
  if (B_past != nullptr && a->sticky) {
    LTM_REQUIRE(u != nullptr, "rect_step: sticky re-sampling needs the uniform draws");
    LTM_PROF(2);
    rc = ltm_resample(a->hist_part, a->H * qtiles, LTM_STICKY_EDGES - 2, 1, u, a->bins, a->bin2basis, 0, a->p,
                      a->b_draw, nullptr, a->ts, a->idx, a->Bv, a->S, stream);
    if (rc) return rc;
    LTM_PROF(3);
    idx = a->idx;
    idx_stride = a->S;
  }
  // tensor-core attention: row-major tf32-rounded K|V from the projection, both attention contractions as UMMAs
  const bool tc512 = a->attn_part != nullptr && a->scores != nullptr && ltm_attn_tc_split_supported(a->N, a->d);
  const bool tcp = a->X != nullptr && a->KV != nullptr && a->precision == 1 && a->gemm_impl == 0 &&
                   (ltm_attn_tc_supported(a->N, a->d) || tc512);
  const bool half_ops = tcp && a->B_half != nullptr && a->Wkv_half != nullptr && a->e % 8 == 0;
  const bool fast = !tcp && a->Kt != nullptr && a->V != nullptr && ltm_attn_fast_supported(a->N, a->d);
  // projected-memory state: every video updates, K|V of the previous call are at hand, and the rows that receive new
  // frames tile the 128-row GEMM tiles evenly -> only those rows are projected (row-major K|V layouts only)
  // (the projected rows start at jg <= jf, chosen so that N - jg divides -- or is a multiple of -- the 128-row GEMM
  // tile: bins in [jg, jf) hold re-sampled memory only and could go either way)
  int n_new = a->N - a->jf;
  if (n_new > 0 && n_new <= 128) {
    int p2 = 1;
    while (p2 < n_new) p2 <<= 1;
    n_new = p2;
  } else if (n_new > 128) {
    n_new = (n_new + 127) / 128 * 128;
  }
  const int jg = a->N - n_new;
  const bool kvstate = a->KV_past != nullptr && a->KV != nullptr && !fast && B_past != nullptr &&
                       new_doc == nullptr && a->jf > 0 && n_new > 0 && jg > 0 && a->e % 4 == 0;
  const int pprec = a->proj_precision ? a->proj_precision : a->precision;
  const bool kvh = a->kv_half != 0 && tcp && a->X16 != nullptr;                    // K|V stored as fp16
  // Blocks of videos run consolidate -> project -> attention back to back, so that the K|V (and coefficient) rows a
  // block has just written are still in L2 when its attention (projection) reads them: with all videos per kernel
  // the 201 MB of K|V at 128 videos go out to HBM and come back.  video_block: 0 = all videos in one block.
  int vblock = (a->video_block > 0 && a->video_block < a->Bv && !fast) ? a->video_block : a->Bv;
  const size_t sB = (size_t)a->N * a->e, sKV = (size_t)a->N * 2 * D, sX = (size_t)xL * xsplits * a->e;
  const bool blocked = vblock < a->Bv;
  if (blocked) LTM_PROF(4);
  for (int v0 = 0; v0 < a->Bv; v0 += vblock) {
    const int nv = (a->Bv - v0 < vblock) ? a->Bv - v0 : vblock;
    const float* Bp = B_past ? B_past + v0 * sB : nullptr;
    float* Bn = a->B_new + v0 * sB;
    // (fp16 K|V: the same pointers, offsets counted in 2-byte elements)
    const size_t kvo = kvh ? v0 * sKV / 2 : v0 * sKV;
    float* KVn = a->KV ? a->KV + kvo : nullptr;
    const float* KVp = kvstate ? a->KV_past + kvo : nullptr;
    void* Bh = half_ops ? (void*)((uint16_t*)a->B_half + v0 * sB) : nullptr;
    if (!blocked) LTM_PROF(4);
    rc = ltm_consolidate_rect_kv(Bp, a->xpart + v0 * sX, idx_stride ? idx + (size_t)v0 * idx_stride : idx, idx_stride,
                                 new_doc ? new_doc + v0 : nullptr, a->seg_ptr0, a->seg_mem0, a->g0, seg_ptr1,
                                 seg_mem1, a->g1, Bn, Bh, KVp, KVn, a->bkv, 2 * D, jg, kvh ? 2 : (tcp ? 1 : 0), nv, a->N,
                                 a->e, xL, xsplits, a->S, stream);
    if (rc) return rc;
    if (!blocked) { LTM_PROF(5); LTM_PROF(6); }
    if (kvstate) {
      // rows [jf, N) of every video as one flat problem: A = B_new + jf * e grouped per video, C = KV + jf * 2D
      ltm_gemm_args ga;
      memset(&ga, 0, sizeof(ga));
      // (fp16 operands: the fp16 copy of the new coefficient rows and of the weights, kind::f16 UMMAs)
      ga.A = half_ops ? reinterpret_cast<const float*>(reinterpret_cast<const uint16_t*>(Bh) + (size_t)jg * a->e)
                      : Bn + (size_t)jg * a->e;
      ga.lda = a->e; ga.a_kmajor = 1;
      ga.a_group = n_new; ga.a_group_stride = (int64_t)a->N * a->e;
      ga.B = half_ops ? reinterpret_cast<const float*>(a->Wkv_half) : a->Wkv; ga.ldb = a->e; ga.b_kmajor = 1;
      ga.ab_fp16 = half_ops ? 1 : 0;
      ga.K1 = a->e; ga.bias = a->bkv;
      ga.C = kvh ? reinterpret_cast<float*>(reinterpret_cast<uint16_t*>(KVn) + (size_t)jg * 2 * D)
                 : KVn + (size_t)jg * 2 * D;
      ga.ldc = 2 * D;
      ga.c_group = n_new; ga.c_group_stride = (int64_t)a->N * 2 * D;
      ga.M = nv * n_new; ga.Nc = 2 * D; ga.K = a->e; ga.batch = 1;
      ga.precision = half_ops ? 1 : pprec; ga.impl = a->gemm_impl; ga.round_tf32 = (tcp && !kvh) ? 1 : 0;
      ga.c_fp16 = kvh ? 1 : 0;
      ga.max_ctas = a->gemm_ctas;
      rc = ltm_gemm(&ga, stream);
    } else if (half_ops) {
      ltm_gemm_args ga;
      memset(&ga, 0, sizeof(ga));
      ga.A = reinterpret_cast<const float*>(Bh); ga.lda = a->e; ga.a_kmajor = 1;
      ga.B = reinterpret_cast<const float*>(a->Wkv_half); ga.ldb = a->e; ga.b_kmajor = 1;
      ga.K1 = a->e; ga.bias = a->bkv;
      ga.C = KVn; ga.ldc = 2 * D;
      ga.M = nv * a->N; ga.Nc = 2 * D; ga.K = a->e; ga.batch = 1;
      ga.precision = 1; ga.impl = 0; ga.round_tf32 = kvh ? 0 : 1; ga.ab_fp16 = 1; ga.c_fp16 = kvh ? 1 : 0;
      ga.max_ctas = a->gemm_ctas;
      rc = ltm_gemm(&ga, stream);
    } else if (kvh) {
      ltm_gemm_args ga;
      memset(&ga, 0, sizeof(ga));
      ga.A = Bn; ga.lda = a->e; ga.a_kmajor = 1;
      ga.B = a->Wkv; ga.ldb = a->e; ga.b_kmajor = 1;
      ga.K1 = a->e; ga.bias = a->bkv;
      ga.C = KVn; ga.ldc = 2 * D;
      ga.M = nv * a->N; ga.Nc = 2 * D; ga.K = a->e; ga.batch = 1;
      ga.precision = pprec; ga.impl = a->gemm_impl; ga.c_fp16 = 1;
      ga.max_ctas = a->gemm_ctas;
      rc = ltm_gemm(&ga, stream);
    } else if (tcp)
      rc = ltm_project_kv_r(Bn, a->Wkv, a->bkv, KVn, nv * a->N, a->e, 2 * D, pprec, a->gemm_impl, stream);
    else if (fast)
      rc = ltm_project_kv_t(Bn, a->Wkv, a->bkv, a->Kt, a->V, nv * a->N, a->e, D, a->N, a->precision, a->gemm_impl,
                            stream);
    else
      rc = ltm_project_kv(Bn, a->Wkv, a->bkv, KVn, nv * a->N, a->e, 2 * D, a->precision, a->gemm_impl, stream);
    if (rc) return rc;
    if (!blocked) { LTM_PROF(7); LTM_PROF(8); }
    const float* qb = q + (size_t)v0 * a->Q * D;
    float* cb = ctx + (size_t)v0 * a->Q * D;
    float* sc = a->scores ? a->scores + (size_t)v0 * a->H * a->Q * a->N : nullptr;
    float* hp = a->sticky ? a->hist_part + (size_t)v0 * a->H * qtiles * (LTM_STICKY_EDGES - 2) : nullptr;
    if (kvh) {
      const uint16_t* Kh = reinterpret_cast<const uint16_t*>(KVn);
      if (tc512)
        rc = ltm_cont_attn_rect_tc16_split(qb, Kh, Kh + D, 2 * D, a->X16, a->W, a->W_out, a->jb, a->tb, cb, sc,
                                           a->attn_part + ltm_attn_tc_split_workspace_floats(v0, a->Q, a->H), hp, nv,
                                           a->Q, a->N, a->H, a->d, stream);
      else
        rc = ltm_cont_attn_rect_tc16(qb, Kh, Kh + D, 2 * D, a->X16, a->W, a->W_out, a->c_none, a->jb, a->tb, cb, sc, hp,
                                     nv, a->Q, a->N, a->H, a->d, stream);
    } else if (tcp && tc512)
      rc = ltm_cont_attn_rect_tc_split(qb, KVn, KVn + D, 2 * D, a->X, a->W, a->W_out, a->jb, a->tb, cb, sc,
                                       a->attn_part + ltm_attn_tc_split_workspace_floats(v0, a->Q, a->H), hp, nv, a->Q,
                                       a->N, a->H, a->d, stream);
    else if (tcp)
      rc = ltm_cont_attn_rect_tc(qb, KVn, KVn + D, 2 * D, a->X, a->W, a->W_out, a->c_none, a->jb, a->tb, cb, sc, hp,
                                 nv, a->Q, a->N, a->H, a->d, stream);
    else if (fast)
      rc = ltm_cont_attn_rect_t(qb, a->Kt, a->V, D, a->W, a->W_out, a->jb, a->tb, cb, sc, hp, nv, a->Q, a->N, a->H,
                                a->d, stream);
    else
      rc = ltm_cont_attn_rect(qb, KVn, a->W, a->W_out, a->jb, a->tb, cb, sc, hp, nv, a->Q, a->N, a->H, a->d, stream);
    if (rc) return rc;
    if (!blocked) LTM_PROF(9);
  }
  if (blocked) LTM_PROF(9);
#undef LTM_PROF
  return rc;
}

#ifdef LTM_BRINGUP
// bring-up: a persistent kernel with a chosen footprint (threads, <= 96 registers, dynamic shared memory) that only
// spins for `us` microseconds: measures what a co-resident footprint alone costs the streaming kernels beside it
__global__ void __maxnreg__(96) footprint_spin_kernel(unsigned long long ns, float* sink) {
  extern __shared__ float fs[];
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (threadIdx.x == 0) fs[0] = 1.f;
  do {
    __nanosleep(200);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < ns);
  if (sink != nullptr && threadIdx.x == 0 && fs[0] == 2.f) sink[0] = 1.f;
}
extern "C" int ltm_debug_footprint_spin(int ctas, int threads, int smem_bytes, int us, void* stream) {
  using namespace ltm;
  LTM_CUDA(cudaFuncSetAttribute(footprint_spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  footprint_spin_kernel<<<ctas, threads, smem_bytes, (cudaStream_t)stream>>>((unsigned long long)us * 1000ull, nullptr);
  LTM_CHECK_LAUNCH("footprint_spin");
  return 0;
}

// bring-up: host time spent in the sections of ltm_rect_step_overlap (sum and max, seconds)
static double g_ov_sum[6] = {0, 0, 0, 0, 0, 0}, g_ov_max[6] = {0, 0, 0, 0, 0, 0};
static inline double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
extern "C" void ltm_debug_overlap_times(double* sum6, double* max6, int reset) {
  for (int i = 0; i < 6; ++i) {
    if (sum6) sum6[i] = g_ov_sum[i];
    if (max6) max6[i] = g_ov_max[i];
    if (reset) g_ov_sum[i] = g_ov_max[i] = 0.0;
  }
}
#define LTM_OV_MARK(i) do { const double t_ = now_s(); const double d_ = t_ - t_prev; g_ov_sum[i] += d_; \
                            if (d_ > g_ov_max[i]) g_ov_max[i] = d_; t_prev = t_; } while (0)
#define LTM_OV_START() double t_prev = now_s()
#else
#define LTM_OV_MARK(i) do { } while (0)
#define LTM_OV_START() do { } while (0)
#endif

extern "C" int ltm_rect_step_overlap(const ltm_rect_step_args* a, const ltm_overlap* o, const float* q,
                                     const double* u, const uint8_t* new_doc, float* ctx) {
  using namespace ltm;
  LTM_REQUIRE(a && o && q && ctx, "rect_step_overlap: null pointer");
  LTM_REQUIRE(o->ev_fork && o->ev_join, "rect_step_overlap: fork / join events missing");
  cudaStream_t ms = (cudaStream_t)o->main_stream, ss = (cudaStream_t)o->side_stream,
               cs = (cudaStream_t)o->compute_stream;
  LTM_OV_START();
  if (o->k_next != nullptr) {
    LTM_REQUIRE(o->xpart_next && o->ev_fork_pool && o->ev_pooled_next, "rect_step_overlap: prefetch target missing");
    // the target buffer was last read by work already queued on the main stream
    LTM_CUDA(cudaEventRecord((cudaEvent_t)o->ev_fork_pool, ms));
    LTM_CUDA(cudaStreamWaitEvent(ss, (cudaEvent_t)o->ev_fork_pool, 0));
    LTM_OV_MARK(0);
    if (a->prof_events[0]) cudaEventRecord((cudaEvent_t)a->prof_events[0], ss);
    int rc = o->next_binned
                 ? ltm_pool_bins(o->k_next, o->xpart_next, a->fbin_ptr, a->Bv, a->L, a->T, a->e, a->xb_rows, ss)
                 : ltm_pool_mean_grid(o->k_next, o->xpart_next, a->Bv, a->L, a->T, a->e, a->splits, o->pool_ctas, ss);
    if (rc) return rc;
    if (a->prof_events[1]) cudaEventRecord((cudaEvent_t)a->prof_events[1], ss);
    LTM_OV_MARK(1);
    LTM_CUDA(cudaEventRecord((cudaEvent_t)o->ev_pooled_next, ss));
    LTM_OV_MARK(2);
  }
  LTM_CUDA(cudaEventRecord((cudaEvent_t)o->ev_fork, ms));
  LTM_CUDA(cudaStreamWaitEvent(cs, (cudaEvent_t)o->ev_fork, 0));
  if (o->ev_pooled_cur) LTM_CUDA(cudaStreamWaitEvent(cs, (cudaEvent_t)o->ev_pooled_cur, 0));
  LTM_OV_MARK(3);
  int rc = ltm_rect_step(a, nullptr, q, u, new_doc, ctx, cs);
  if (rc) return rc;
  LTM_OV_MARK(4);
  LTM_CUDA(cudaEventRecord((cudaEvent_t)o->ev_join, cs));
  LTM_CUDA(cudaStreamWaitEvent(ms, (cudaEvent_t)o->ev_join, 0));
  LTM_OV_MARK(5);
  return 0;
}

extern "C" int ltm_event_create(void** ev) {
  using namespace ltm;
  LTM_REQUIRE(ev != nullptr, "event_create: null pointer");
  cudaEvent_t e;
  LTM_CUDA(cudaEventCreate(&e));
  *ev = (void*)e;
  return 0;
}
extern "C" int ltm_event_create_sync(void** ev) {
  using namespace ltm;
  LTM_REQUIRE(ev != nullptr, "event_create_sync: null pointer");
  cudaEvent_t e;
  LTM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *ev = (void*)e;
  return 0;
}
extern "C" int ltm_stream_wait_event(void* stream, void* ev) {
  using namespace ltm;
  LTM_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0));
  return 0;
}
extern "C" int ltm_event_record(void* ev, void* stream) {
  using namespace ltm;
  LTM_CUDA(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream));
  return 0;
}
extern "C" int ltm_event_elapsed_ms(void* start, void* stop, float* ms) {
  using namespace ltm;
  LTM_REQUIRE(ms != nullptr, "event_elapsed_ms: null pointer");
  cudaError_t e = cudaEventSynchronize((cudaEvent_t)stop);
  if (e == cudaSuccess) e = cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();            // e.g. an event that was never recorded: do not leave a stale error behind
    set_error("event_elapsed_ms: %s", cudaGetErrorString(e));
    return -3;
  }
  return 0;
}
extern "C" int ltm_event_destroy(void* ev) {
  using namespace ltm;
  LTM_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return 0;
}

extern "C" int ltm_rect_step_host(const ltm_rect_step_args* a, const float* k_host, const float* q_host,
                                  const double* u_host, const uint8_t* new_doc_host, float* ctx_host, void* stream) {
  using namespace ltm;
  LTM_REQUIRE(a && k_host && q_host && ctx_host, "rect_step_host: null pointer");
  LTM_REQUIRE(a->k_dev && a->q_dev && a->ctx_dev, "rect_step_host: device staging buffers missing");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t D = (size_t)a->H * a->d;
  LTM_CUDA(cudaMemcpyAsync(a->k_dev, k_host, sizeof(float) * (size_t)a->Bv * a->L * a->T * a->e,
                           cudaMemcpyHostToDevice, st));
  LTM_CUDA(cudaMemcpyAsync(a->q_dev, q_host, sizeof(float) * (size_t)a->Bv * a->Q * D, cudaMemcpyHostToDevice, st));
  const double* u_dev = nullptr;
  if (u_host) {
    LTM_REQUIRE(a->u_dev, "rect_step_host: u staging buffer missing");
    LTM_CUDA(cudaMemcpyAsync(a->u_dev, u_host, sizeof(double) * (size_t)a->Bv * a->S, cudaMemcpyHostToDevice, st));
    u_dev = a->u_dev;
  }
  const uint8_t* nd_dev = nullptr;
  if (new_doc_host) {
    LTM_REQUIRE(a->new_doc_dev, "rect_step_host: new_doc staging buffer missing");
    LTM_CUDA(cudaMemcpyAsync(a->new_doc_dev, new_doc_host, (size_t)a->Bv, cudaMemcpyHostToDevice, st));
    nd_dev = a->new_doc_dev;
  }
  int rc = ltm_rect_step(a, a->k_dev, a->q_dev, u_dev, nd_dev, a->ctx_dev, stream);
  if (rc) return rc;
  LTM_CUDA(cudaMemcpyAsync(ctx_host, a->ctx_dev, sizeof(float) * (size_t)a->Bv * a->Q * D, cudaMemcpyDeviceToHost, st));
  return 0;
}
