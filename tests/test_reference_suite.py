"""Drop-in check: the reference's OWN test suite (/root/reference/tests, 40 tests of src/utils.py) run unmodified
against this repository's ``src`` package.  Authoring container only - the reference does not travel to the GPU
box - and read-only: no cache, no bytecode is written next to the reference's files."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

RUNNER = r'''
import sys, types
# the reference imports google.cloud.storage at module import (src/utils.py:8); the package is not installed here
g, gc, gs = types.ModuleType("google"), types.ModuleType("google.cloud"), types.ModuleType("google.cloud.storage")
gs.Client = type("Client", (), {}); gc.storage = gs; g.cloud = gc
sys.modules.update({"google": g, "google.cloud": gc, "google.cloud.storage": gs})
sys.path.insert(0, sys.argv[1])          # `src` resolves to this repository's package, not the reference's
sys.dont_write_bytecode = True
import pytest
sys.exit(pytest.main(["/root/reference/tests", "-p", "no:cacheprovider", "--import-mode=importlib", "-q",
                      "--rootdir=" + sys.argv[1]]))
'''


@pytest.mark.reference
def test_reference_test_suite_passes_against_this_package():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", RUNNER, str(ROOT)], capture_output=True, text=True, env=env,
                       cwd=str(ROOT), timeout=600)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
