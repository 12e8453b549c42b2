"""The drop-in boundary is a C ABI: a plain C99 client (no C++, no Python, no torch types) must compile against
include/dracob200.h, link against libdracob200.so and walk the reference's sample asset with it -- host phases only
(dcb_index with a NULL context, dcb_host_connectivity, dcb_index_finish, the getters), so no GPU is needed.  This is the
binding a [DllImport] / cgo / JNI stub makes, written out in C."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(ROOT, "draco_sharp_b200")

CLIENT = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dracob200.h"

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 3;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t *buf = (uint8_t *)malloc((size_t)n);
  if (fread(buf, 1, (size_t)n, f) != (size_t)n) return 4;
  fclose(f);
  const uint8_t *bufs[1];
  uint64_t lens[1];
  dcb_batch *b = NULL;
  dcb_buffer_info bi;
  dcb_attr_info ai;
  int rc, a;
  bufs[0] = buf;
  lens[0] = (uint64_t)n;
  printf("version %d\n", dcb_version());
  rc = dcb_index(NULL, bufs, lens, 1, &b);                 /* NULL ctx: host indexing only */
  if (rc) { printf("index rc %d\n", rc); return 5; }
  rc = dcb_host_connectivity(b, 0);                         /* Edgebreaker connectivity on the host */
  if (rc) { printf("connectivity rc %d\n", rc); return 6; }
  rc = dcb_index_finish(NULL, b);
  if (rc) { printf("finish rc %d\n", rc); return 7; }
  memset(&bi, 0, sizeof bi);
  rc = dcb_get_buffer_info(b, 0, &bi);
  if (rc) return 8;
  printf("status %d points %u attrs %d\n", (int)bi.status, (unsigned)bi.n_points, (int)bi.n_attrs);
  for (a = 0; a < bi.n_attrs; ++a) {
    memset(&ai, 0, sizeof ai);
    if (dcb_get_attr_info(b, 0, a, &ai)) return 9;
    printf("attr %d pred %d entries %u bytes %llu\n", a, (int)ai.pred_method, (unsigned)ai.n_entries, (unsigned long long)ai.out_bytes);
  }
  dcb_batch_free(b);
  free(buf);
  return 0;
}
'''


def test_a_c99_client_compiles_links_and_indexes_the_sample(tmp_path):
    lib = os.path.join(LIB_DIR, "libdracob200.so")
    if not os.path.exists(lib):
        pytest.skip("libdracob200.so not built")
    src = tmp_path / "client.c"
    src.write_text(CLIENT)
    exe = tmp_path / "client"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", LIB_DIR, "-ldracob200", "-Wl,-rpath," + LIB_DIR])
    out = subprocess.run([str(exe), os.path.join(ROOT, "tests", "golden", "house_04.obj.drc")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[1] == "status 0 points 3220 attrs 3"
    assert lines[2].startswith("attr 0 pred 1 entries 1775 bytes 21300")
    assert lines[3].startswith("attr 1 pred 5 entries 3220 bytes 25760")
