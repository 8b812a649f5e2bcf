"""Build oracle/_ref/ from the UNMODIFIED reference under /root/reference  --  TEST / BASELINE INFRASTRUCTURE.

The reference is pure Python, so "building" it means byte-compiling the three modules of the hot path
(flow_realnvp.py, modules_realnvp.py, utils.py) where they lie into sourceless byte-code files (``<module>.bin``: the snapshot that ships the repo to the GPU box skips ``*.pyc``) under
``oracle/_ref/``.  No reference source is copied into the repository: ``oracle/_ref/`` is git-ignored, holds
build outputs only and travels to the GPU box with the snapshot, exactly like the repo's own ``.so`` files.  There
``bench.py --impl reference`` / ``cpu_baseline`` / ``gpu_eager_baseline`` import it (``oracle/ref_loader.py``) and
time the reference's own code path (``kind: "reference"``); without it they fall back to the oracle port
(``kind: "port"``).

Run:  python oracle/build_ref.py            (needs /root/reference; done by __graft_entry__.build())
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RNVP_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ("flow_realnvp", "modules_realnvp", "utils")


def build() -> bool:
    if not all(os.path.exists(os.path.join(REF, m + ".py")) for m in MODULES):
        return False
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        # dfile: the path recorded in tracebacks stays the reference's own
        py_compile.compile(os.path.join(REF, m + ".py"), cfile=os.path.join(OUT, m + ".bin"),
                           dfile=f"reference/{m}.py", doraise=True, optimize=0)
    with open(os.path.join(OUT, "BUILD_INFO"), "w") as f:
        f.write(f"byte-compiled from {REF} by oracle/build_ref.py with python {sys.version.split()[0]}\n")
    # a checkpoint pair written by the reference's own torch.save calls (interop fixture, oracle/make_golden_api.py)
    if not os.path.exists(os.path.join(OUT, "ckpt", "expect.pt")):
        import subprocess
        subprocess.run([sys.executable, os.path.join(HERE, "make_golden_api.py"), "--checkpoint"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return True


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref built" if ok else f"{REF} not present: oracle/_ref not built")
