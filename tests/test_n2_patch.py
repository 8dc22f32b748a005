"""the GPU colour stage reachable through the reference's public API.  oracle/_ref/libheif_ref_b200.so is the
reference core plus integration/b200_color_op.cc (one ColorConversionOperation with SpeedCosts_Hardware backed by
libb200heif.so) and the three-line integration/colorconversion_b200.patch; heif_decode_image() through it must return the
same bytes as through the unmodified build for BASELINE configs 1-4 (scaled down)."""
import json
import os
import subprocess
import sys

import pytest

from oracle import bindings as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
have = all(os.path.exists(os.path.join(ob.REF, f)) for f in ("libheif_ref.so", "libheif_ref_b200.so", "liboracle_plugin.so", "liboracle_plugin_b200.so")) and ob.avcodec_dir()


def run(lib):
    env = dict(os.environ, B200_REF_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "n2_child.py")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="oracle/_ref builds (libheif_ref.so, libheif_ref_b200.so) not present")
def test_patched_libheif_decodes_identically(cuda):
    ref = run("libheif_ref.so")
    gpu = run("libheif_ref_b200.so")
    assert ref.keys() == gpu.keys() and len(ref) >= 7
    for k in ref:
        assert ref[k] == gpu[k], k
