"""Run the REFERENCE's own test files against this package's host-side mirror (build container only: needs /root/reference).

The reference tests are copied to a scratch directory, their conftest is pointed at `birdnet-stm32_b200/` instead of the
reference package, and bare stand-ins for `tensorflow` / `soundfile` let the modules that only guard themselves with
`pytest.importorskip("tensorflow")` be collected.  Nothing from the reference package is imported.  usage:
    python scripts/run_reference_tests.py [test_file_stem ...]
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/tests"
DEFAULT = ["test_pooling", "test_threshold_opt", "test_frontend_registry", "test_metrics"]

STUBS = '''sys.path.insert(0, %r)
import types
if "tensorflow" not in sys.modules:
    sys.modules["tensorflow"] = types.ModuleType("tensorflow")
if "soundfile" not in sys.modules:
    _sf = types.ModuleType("soundfile")
    def _write(path, audio, sr, subtype=None):
        from birdnet_stm32.audio.io import save_wav
        save_wav(np.asarray(audio), str(path), int(sr))
    _sf.write = _write
    sys.modules["soundfile"] = _sf
'''


def main():
    if not os.path.isdir(REF_TESTS):
        print("reference checkout not mounted")
        return 0
    names = sys.argv[1:] or DEFAULT
    tmp = tempfile.mkdtemp(prefix="reftests_")
    for f in os.listdir(REF_TESTS):
        if f.endswith(".py"):
            shutil.copy(os.path.join(REF_TESTS, f), tmp)
    conf = open(os.path.join(tmp, "conftest.py")).read()
    conf = conf.replace('sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))', STUBS % os.path.join(ROOT, "birdnet-stm32_b200"))
    open(os.path.join(tmp, "conftest.py"), "w").write(conf)
    rc = 0
    for n in names:
        r = subprocess.run([sys.executable, "-m", "pytest", f"{n}.py", "-q", "-p", "no:cacheprovider"], cwd=tmp, capture_output=True, text=True)
        print(f"{n}: {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip()[-200:]}")
        rc |= r.returncode
    shutil.rmtree(tmp, ignore_errors=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())          # pytest's own code: 0 = all passed, 1 = some reference tests failed
