/* A non-Python embedder of librho_b200: plain C, only include/rho_b200.h and the C runtime (no torch, no CUDA headers).
 * Reads n fixed-length clips from a raw fp32 file, runs rho_b200_validate_host (host buffers in, host buffers out) and
 * writes the records, the processed audio and the feature rows to raw files, which tests/test_c_embedder.py compares
 * with what the Python host mirror produces for the same clips.
 *
 *   gcc -O2 -I include tests/c/host_embedder.c -o /tmp/host_embedder -L rho_tts_b200 -lrho_b200 -Wl,-rpath,$PWD/rho_tts_b200
 *   /tmp/host_embedder clips.f32 n clip_len out_prefix
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rho_b200.h"

static int die(const char* what) {
  fprintf(stderr, "host_embedder: %s: %s\n", what, rho_b200_last_error());
  return 1;
}

static int dump(const char* prefix, const char* suffix, const void* p, size_t bytes) {
  char path[1024];
  snprintf(path, sizeof(path), "%s.%s", prefix, suffix);
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  const size_t w = fwrite(p, 1, bytes, f);
  fclose(f);
  return w != bytes;
}

int main(int argc, char** argv) {
  if (argc != 5) { fprintf(stderr, "usage: host_embedder clips.f32 n clip_len out_prefix\n"); return 2; }
  const int n = atoi(argv[2]);
  const int clip_len = atoi(argv[3]);
  const size_t samples = (size_t)n * (size_t)clip_len;
  float* x = (float*)malloc(samples * sizeof(float));
  float* y = (float*)malloc(samples * sizeof(float));
  float* mel = (float*)malloc((size_t)n * 80 * 3000 * sizeof(float));
  rho_record* rec = (rho_record*)calloc((size_t)n, sizeof(rho_record));
  if (!x || !y || !mel || !rec) return 3;
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(x, sizeof(float), samples, f) != samples) { fprintf(stderr, "host_embedder: cannot read %s\n", argv[1]); return 3; }
  fclose(f);

  if (rho_b200_abi_version() != RHO_B200_ABI_VERSION) { fprintf(stderr, "host_embedder: ABI mismatch\n"); return 4; }
  rho_handle* h = NULL;
  if (rho_b200_create(&h, 0) != RHO_OK) return die("rho_b200_create");
  /* the BaseTTS defaults (base_tts.py:72-81): sr, trim on, -50 dB, 20 ms fades, 50 ms crossfade, 100 ms pause, decay 0.3 */
  rho_params p = {24000, 1, -50.0, 0.02, 0.05, 0.1, 0.3};
  /* pageable host memory works too (slower than pinned: the driver stages it) */
  if (rho_b200_validate_host(h, x, n, clip_len, &p, y, 80, 3000, mel, NULL, NULL, 0, rec) != RHO_OK)
    return die("rho_b200_validate_host");
  int accepted = 0;
  for (int i = 0; i < n; ++i) accepted += rec[i].ok;
  printf("host_embedder: %d clips of %d samples, %d accepted, %lld kernel launches\n", n, clip_len, accepted,
         (long long)rho_b200_launch_count(h));
  if (dump(argv[4], "rec", rec, (size_t)n * sizeof(rho_record)) || dump(argv[4], "y", y, samples * sizeof(float)) ||
      dump(argv[4], "mel", mel, (size_t)n * 80 * 3000 * sizeof(float)))
    return 5;
  rho_b200_destroy(h);
  free(x); free(y); free(mel); free(rec);
  return 0;
}
