"""TEST / BENCH INFRASTRUCTURE -- not product code.

Stages an UNMODIFIED copy of the reference's hot-path Python modules (the files SURVEY.md section 8a cites: the CFM
solve, the estimator, the matcha transformer block, the DAC-VAE decoder) from /root/reference into ``baseline/_ref/``,
which is git-ignored (no reference source enters the history) but travels to the GPU box with the snapshot
(operational note iii).  ``bench.py --impl reference`` and the ``cpu_baseline`` /
``gpu_eager`` legs then time the reference's OWN modules through its own call surface
(``CausalConditionalCFM.forward`` -> ``DACVAE.decode``) instead of the restatement in oracle/restatement.py; the
third-party packages the reference imports and this image lacks are the stubs of oracle/ref_import.py.

The file list is not hard-coded: the reference is imported here (build container) through oracle/ref_import.py and
every module that was loaded from /root/reference is copied byte for byte, with a sha256 manifest.

    python -m oracle.stage_ref          (also run by __graft_entry__.build() when /root/reference exists)
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "speech", "cosyvoice")):
        return None  # not the build container: whatever was staged earlier is used as it is
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["LS_REFERENCE_ROOT"] = SRC
    from oracle import ref_import as R
    R.build_reference_flow()
    R.build_reference_noncausal_estimator()
    R.build_reference_dac()
    files = sorted({m.__file__ for m in list(sys.modules.values())
                    if getattr(m, "__file__", None) and m.__file__.startswith(SRC + os.sep)})
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for f in files:
        rel = os.path.relpath(f, SRC)
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(f, out)
        with open(out, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(files)} unmodified reference modules into {DST}")
    return DST


if __name__ == "__main__":
    stage()
