"""The ctypes mirrors of the C-ABI structs (binaural-audio-synthesis_b200/_cabi.py) against the header itself: a small C
program that includes include/bas_b200.h prints sizeof / offsetof of every struct member the Python side fills in; the
numbers must match ctypes'.  A field added on one side only would otherwise shift every later pointer silently."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_layout(tmp_path, structs):
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "bas_b200.h"', 'int main(void) {']
    for name, fields in structs.items():
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f in fields:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return {k: int(v) for k, v in (line.split() for line in out.splitlines())}


def test_ctypes_structs_match_the_header(tmp_path):
    sys.path.insert(0, ROOT)
    import binaural_audio_synthesis_b200 as bas
    cabi = bas._cabi
    mirrors = {'bas_route': cabi.Route, 'bas_step_job': cabi.StepJob, 'bas_pipeline_job': cabi.PipelineJob, 'bas_term': cabi.Term}
    structs = {name: [f[0] for f in cls._fields_] for name, cls in mirrors.items()}
    got = _c_layout(tmp_path, structs)
    for name, cls in mirrors.items():
        assert got[name] == C.sizeof(cls), name
        for field, *_ in cls._fields_:
            assert got['%s.%s' % (name, field)] == getattr(cls, field).offset, (name, field)
