/* A plain-C consumer of the C ABI (include/rlrm_b200.h): no Python, no torch. It reads a scenario blob written by
 * tables.Compiled (tests/test_c_abi_demo.py: rlrm_config_t bytes followed by the flat tables), allocates the state with the
 * CUDA runtime, runs rlrm_reset + rlrm_train and prints what a caller would look at. This is the binding a non-Python host
 * (the "reference-side stub" of INTEGRATION.md) performs.
 *
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/c_abi_demo.c -o c_abi_demo \
 *       -Lmultiagent-rl-rm_b200 -lrlrm_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/multiagent-rl-rm_b200
 *   ./c_abi_demo scenario.blob <n_instances> <n_iters> q_out.bin
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rlrm_b200.h"

#define CK(x)                                                                 \
  do {                                                                        \
    if ((x) != cudaSuccess) { fprintf(stderr, "CUDA error at %s\n", #x); return 2; } \
  } while (0)
#define RL(x)                                                                                  \
  do {                                                                                         \
    if ((x) != RLRM_OK) { fprintf(stderr, "rlrm error at %s: %s\n", #x, rlrm_last_error()); return 3; } \
  } while (0)

static void* read_array(FILE* f, uint64_t* n_bytes) {
  if (fread(n_bytes, sizeof(*n_bytes), 1, f) != 1) return NULL;
  if (*n_bytes == 0) return NULL;
  void* p = malloc(*n_bytes);
  if (!p || fread(p, 1, *n_bytes, f) != *n_bytes) return NULL;
  return p;
}

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s scenario.blob n_instances n_iters q_out.bin\n", argv[0]); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  rlrm_config_t cfg;
  double q_init;
  if (fread(&cfg, sizeof(cfg), 1, f) != 1 || fread(&q_init, sizeof(q_init), 1, f) != 1) { fprintf(stderr, "short blob\n"); return 1; }
  rlrm_tables_t tb;
  memset(&tb, 0, sizeof(tb));
  uint64_t nb;
  tb.next_cell = read_array(f, &nb);
  tb.cell_flags = read_array(f, &nb);
  tb.label = read_array(f, &nb);
  tb.delta = read_array(f, &nb);
  tb.rq = read_array(f, &nb);
  tb.rcf = read_array(f, &nb);
  tb.qrm_states = read_array(f, &nb);
  tb.start_cell = read_array(f, &nb);
  tb.free_cells = read_array(f, &nb);
  tb.phi = read_array(f, &nb);
  fclose(f);

  const int64_t N = atoll(argv[2]);
  const int n_iters = atoi(argv[3]);
  const int A = cfg.n_agents;
  const size_t S = (size_t)cfg.width * cfg.height * cfg.n_rm_states;
  const size_t n_slots = (size_t)N * A, n_q = n_slots * S * 4;

  if (rlrm_abi_version() != RLRM_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
  rlrm_handle_t* h = NULL;
  RL(rlrm_create(&cfg, &tb, 0, &h));

  rlrm_state_t st;
  memset(&st, 0, sizeof(st));
  st.n_instances = N;
  CK(cudaMalloc((void**)&st.slot, n_slots * sizeof(uint64_t)));
  CK(cudaMalloc((void**)&st.epsilon, n_slots * sizeof(double)));
  CK(cudaMalloc((void**)&st.q, n_q * sizeof(float)));
  CK(cudaMalloc((void**)&st.ep_return, n_slots * sizeof(double)));
  CK(cudaMalloc((void**)&st.stats, n_slots * sizeof(rlrm_stats_t)));
  CK(cudaMemset(st.slot, 0, n_slots * sizeof(uint64_t)));
  CK(cudaMemset(st.ep_return, 0, n_slots * sizeof(double)));
  CK(cudaMemset(st.stats, 0, n_slots * sizeof(rlrm_stats_t)));
  double* eps = malloc(n_slots * sizeof(double));
  float* q = malloc(n_q * sizeof(float));
  for (size_t k = 0; k < n_slots; k++) eps[k] = cfg.epsilon_start;
  for (size_t k = 0; k < n_q; k++) q[k] = (float)q_init;
  CK(cudaMemcpy(st.epsilon, eps, n_slots * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(st.q, q, n_q * sizeof(float), cudaMemcpyHostToDevice));

  RL(rlrm_reset(h, &st, NULL, NULL));                      /* rm_env.reset(seed) before the episode loop ... */
  RL(rlrm_reset(h, &st, NULL, NULL));                      /* ... and at the start of the first episode       */
  RL(rlrm_train(h, &st, 0, n_iters, 1, NULL, NULL));       /* n_iters iterations of the driver loop           */
  CK(cudaDeviceSynchronize());

  rlrm_stats_t* stats = malloc(n_slots * sizeof(rlrm_stats_t));
  CK(cudaMemcpy(stats, st.stats, n_slots * sizeof(rlrm_stats_t), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(q, st.q, n_q * sizeof(float), cudaMemcpyDeviceToHost));
  uint64_t episodes = 0, successes = 0;
  for (size_t k = 0; k < n_slots; k++) { episodes += stats[k].episodes; successes += stats[k].successes; }
  printf("instances=%lld agents=%d iterations=%d episodes=%llu successes=%llu launches=%lld\n", (long long)N, A, n_iters,
         (unsigned long long)episodes, (unsigned long long)successes, (long long)rlrm_launch_count(h));
  FILE* out = fopen(argv[4], "wb");
  if (!out || fwrite(q, sizeof(float), n_q, out) != n_q) { perror(argv[4]); return 1; }
  fclose(out);
  RL(rlrm_destroy(h));
  return 0;
}
