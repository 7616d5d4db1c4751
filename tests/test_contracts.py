"""Repository contracts that need no GPU: the public header is valid C, the bench reference arm
prints the JSON line the driver parses, and the product package never imports the oracle."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / 'use_header.c'
    src.write_text('#include "tgm_b200.h"\n'
                   'int probe(void) { tgm_dyg_params p; (void)p; return TGM_OK + TGM_HOST_SLOTS; }\n')
    out = subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-c', str(src), '-I',
                          os.path.join(ROOT, 'include'), '-o', str(tmp_path / 'use_header.o')],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_c_program_links_the_library_and_fails_loudly_without_a_gpu(tmp_path):
    """A C caller (no Python, no torch) binds the ABI; device entry points report errors through
    tgm_last_error() instead of computing anything on the CPU."""
    src = tmp_path / 'cabi_probe.c'
    src.write_text('''
#include <stdio.h>
#include <string.h>
#include "tgm_b200.h"
int main(void) {
  tgm_recency *r = 0;
  int64_t lb = -1, ub = -1;
  int32_t s[3] = {0, 1, 2}, d[3] = {1, 2, 0};
  int64_t t[3] = {5, 5, 9};
  tgm_store *st = 0;
  if (tgm_version() < 100) return 1;
  if (tgm_store_create(&st, s, d, t, 0, 3, 0, 3, -1, TGM_MEM_HOST, 0) != TGM_OK) return 2;
  if (tgm_store_bounds(st, 5, 1, 5, 1, -1, -1, &lb, &ub) != TGM_OK || lb != 0 || ub != 2) return 3;
  tgm_store_destroy(st);
  if (tgm_device_count() == 0) {
    if (tgm_recency_create(&r, 8, 2, 0, 0) >= 0 || r != 0) return 4;
    if (strlen(tgm_last_error()) == 0) return 5;
  }
  printf("ok\\n");
  return 0;
}
''')
    exe = tmp_path / 'cabi_probe'
    libdir = os.path.join(ROOT, 'tgm_b200', 'csrc')
    build = subprocess.run(['gcc', '-std=c99', str(src), '-I', os.path.join(ROOT, 'include'), '-L', libdir,
                            '-ltgm_b200', f'-Wl,-rpath,{libdir}', '-o', str(exe)],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.strip() == 'ok', (run.returncode, run.stderr)


def test_bench_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '2', '--warmup', '1'], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'sampled-edges/s' and d['value'] > 0
    assert d['higher_is_better'] is True and d['steps'] == 2 and d['warmup'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_product_package_never_imports_the_oracle():
    pat = re.compile(r'^\s*(from|import)\s+oracle\b', re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'tgm_b200')):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f'{f} imports the oracle'
