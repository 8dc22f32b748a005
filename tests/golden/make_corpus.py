"""Extracts every HEVC access unit the UNMODIFIED reference pushes into a decoder plugin while it tries to decode the files
of its fuzzing corpus (fuzzing/data/corpus/*, mostly minimised fuzzer findings) into tests/golden/corpus/*.au, with the same
capture mechanism as make_streams.py (CPU oracle plugin, B200_ORACLE_DUMP_DIR).  Run in the build container:

    python tests/golden/make_corpus.py

The GPU suite then feeds every one of them to the CUDA decoder: "an error code or the oracle's planes, never a crash or a
hang" (tests/test_corpus_gpu.py); the CPU suite does the same with the host front-end."""
import hashlib
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/fuzzing/data/corpus"
OUT = os.path.join(ROOT, "tests", "golden", "corpus")

CHILD = r"""
import sys
sys.path.insert(0, %r)
from oracle import refheif as rh
rh.load(); rh.register_cpu_decoder()
try:
    rh.decode_file(sys.argv[1], decoder_id="b200-oracle", threads=1)
except Exception as e:
    pass
"""


def main():
    os.makedirs(OUT, exist_ok=True)
    seen = {}
    for name in sorted(os.listdir(REF)):
        with tempfile.TemporaryDirectory() as d:
            env = dict(os.environ, B200_ORACLE_DUMP_DIR=d)
            try:
                subprocess.run([sys.executable, "-c", CHILD % ROOT, os.path.join(REF, name)], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=60)
            except subprocess.TimeoutExpired:
                print("timeout", name)
            for f in sorted(os.listdir(d)):
                b = open(os.path.join(d, f), "rb").read()
                if 8 <= len(b) <= 256 * 1024:
                    h = hashlib.md5(b).hexdigest()[:12]
                    if h not in seen:
                        seen[h] = name
                        open(os.path.join(OUT, h + ".au"), "wb").write(b)
    with open(os.path.join(OUT, "INDEX.txt"), "w") as f:
        for h, n in sorted(seen.items()):
            f.write(f"{h}.au  first pushed while decoding fuzzing/data/corpus/{n}\n")
    print(len(seen), "access units")


if __name__ == "__main__":
    main()
