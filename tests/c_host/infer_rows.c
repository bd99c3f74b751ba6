/* A host without Python: the reference's request — model, system prompt, output schema (as the
 * JSON text `normalize_output_schema` produces, sutro/common.py:152-163) and the input rows —
 * served through the C-ABI of libsutro_b200.so alone (include/sutro_b200.h).
 *
 *   infer_rows <bundle_dir> <system_prompt> <schema.json | -> <row> [<row> ...]
 *
 * prints one output per line, in row order (results are positional, sutro/sdk.py:406-412).
 * Build:  gcc -std=c11 -I include tests/c_host/infer_rows.c -L sutro_b200 -lsutro_b200 \
 *             -Wl,-rpath,$PWD/sutro_b200 -o infer_rows                                     */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sutro_b200.h"

static char* slurp(const char* path, long* len) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  *len = ftell(f);
  fseek(f, 0, SEEK_SET);
  char* buf = (char*)malloc(*len + 1);
  if (fread(buf, 1, *len, f) != (size_t)*len) {
    fclose(f);
    free(buf);
    return NULL;
  }
  buf[*len] = 0;
  fclose(f);
  return buf;
}

int main(int argc, char** argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <bundle_dir> <system_prompt> <schema.json|-> <row>...\n", argv[0]);
    return 2;
  }
  const int n_rows = argc - 4;
  long schema_len = 0;
  char* schema = strcmp(argv[3], "-") ? slurp(argv[3], &schema_len) : NULL;
  if (strcmp(argv[3], "-") && !schema) {
    fprintf(stderr, "cannot read %s\n", argv[3]);
    return 2;
  }
  /* Arrow-style column: one byte blob + int64 offsets[n+1] */
  int64_t* off = (int64_t*)calloc(n_rows + 1, sizeof(int64_t));
  size_t total = 0;
  for (int i = 0; i < n_rows; ++i) total += strlen(argv[4 + i]);
  uint8_t* bytes = (uint8_t*)malloc(total ? total : 1);
  for (int i = 0; i < n_rows; ++i) {
    const size_t n = strlen(argv[4 + i]);
    memcpy(bytes + off[i], argv[4 + i], n);
    off[i + 1] = off[i] + (int64_t)n;
  }
  void* model = NULL;
  if (sb200_model_open(argv[1], 0, 8, 512, 512, &model)) {
    fprintf(stderr, "model_open: %s\n", sb200_last_error());
    return 1;
  }
  sb200_result* res = NULL;
  sb200_job_stats st;
  memset(&st, 0, sizeof st);
  if (sb200_model_infer(model, argv[2], schema, schema_len, NULL, 0, NULL, bytes, off, n_rows, 1,
                        &res, &st)) {
    fprintf(stderr, "model_infer: %s\n", sb200_last_error());
    return 1;
  }
  for (int64_t i = 0; i < res->n_rows; ++i) {
    fwrite(res->bytes + res->offsets[i], 1, (size_t)(res->offsets[i + 1] - res->offsets[i]), stdout);
    fputc('\n', stdout);
  }
  fprintf(stderr, "rows %lld input_tokens %lld decode_tokens %lld device_ms %.2f\n",
          (long long)st.rows_done, (long long)st.input_tokens, (long long)st.decode_tokens,
          st.t_device_ms);
  sb200_result_free(res);
  sb200_model_close(model);
  free(bytes);
  free(off);
  free(schema);
  return 0;
}
