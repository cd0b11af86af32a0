// Closed-loop load generator over the public C ABI (include/tfsc_b200.h): `concurrency` client
// threads each issue tfsc_predict() calls with HOST buffers, one request at a time, exactly as a
// cgo Predict handler would. Used by bench.py for the end-to-end (e2e) number and latency
// percentiles. Links against nothing but libtfsc_b200.so's C entry points (resolved at run time
// through function pointers handed in by the caller), so it measures the product path only.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/tfsc_b200.h"

extern "C" {

typedef int (*predict_fn)(tfsc_server*, const char*, const char*, const tfsc_tensor*, int, tfsc_tensor*, int);
typedef int (*predict_member_fn)(tfsc_server*, int, const char*, const char*, const tfsc_tensor*, int, tfsc_tensor*, int, int64_t);

// names: n_models NUL-terminated strings packed back to back with stride name_stride.
// trace[i] = model index of request i. inputs: n_inputs rows of in_dim floats (request i uses row
// i % n_inputs). outputs: concurrency * out_dim floats of scratch. lat_us[i] = latency of request i.
// members (optional): members[i] = index into gpu.members of the cache node request i was routed to by the caller (the
// front tier's decision); such requests go through tfsc_predict_member (a member of another rank takes the forward hop).
// Returns the number of failed requests.
int64_t tfsc_loadgen_run(void* predict, tfsc_server* srv, const char* names, int name_stride, const char* version,
                         const int32_t* trace, int64_t n_requests, const float* inputs, int64_t n_inputs, int in_dim,
                         float* outputs, int out_dim, int concurrency, float* lat_us, double* elapsed_s,
                         void* predict_member, const int32_t* members) {
  predict_fn fn = reinterpret_cast<predict_fn>(predict);
  predict_member_fn fnm = reinterpret_cast<predict_member_fn>(predict_member);
  std::atomic<int64_t> next{0}, failed{0};
  auto t0 = std::chrono::steady_clock::now();
  auto worker = [&](int tid) {
    float* y = outputs + (size_t)tid * out_dim;
    for (;;) {
      int64_t i = next.fetch_add(1);
      if (i >= n_requests) break;
      tfsc_tensor in;
      memset(&in, 0, sizeof in);
      in.dtype = TFSC_DT_FLOAT;
      in.rank = 2;
      in.shape[0] = 1;
      in.shape[1] = in_dim;
      in.data = const_cast<float*>(inputs + (size_t)(i % n_inputs) * in_dim);
      in.nbytes = (size_t)in_dim * 4;
      tfsc_tensor out;
      memset(&out, 0, sizeof out);
      out.data = y;
      out.nbytes = (size_t)out_dim * 4;
      auto a = std::chrono::steady_clock::now();
      int rc = (fnm && members) ? fnm(srv, members[i], names + (size_t)trace[i] * name_stride, version, &in, 1, &out, 1, 0)
                                : fn(srv, names + (size_t)trace[i] * name_stride, version, &in, 1, &out, 1);
      auto b = std::chrono::steady_clock::now();
      if (rc < 0) failed++;
      if (lat_us) lat_us[i] = (float)std::chrono::duration<double, std::micro>(b - a).count();
    }
  };
  std::vector<std::thread> ts;
  for (int t = 0; t < concurrency; ++t) ts.emplace_back(worker, t);
  for (auto& t : ts) t.join();
  if (elapsed_s) *elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return failed.load();
}
}
