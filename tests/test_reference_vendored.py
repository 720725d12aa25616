"""The UNMODIFIED reference, staged under oracle/_ref/ by tools/vendor_ref.sh (byte copies; git-ignored, travels to the
GPU box with the snapshot).

CPU: the staged files are what SHA256SUMS says, and - in the build container - what /root/reference holds.
GPU: the reference's OWN test files (tests/test_ficp.py:39-126, tests/test_rigid_2d_operations.py:17-75) run VERBATIM in a
subprocess whose `from ficp import FractionalICP` resolves to THIS repo's drop-in (ficp.py -> coregistrationgame_b200);
`trees` resolves to the reference's own domain model, the consumer of the result.
"""
import hashlib
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "oracle", "_ref")
FILES = ["ficp.py", "trees.py", "tests/test_ficp.py", "tests/test_rigid_2d_operations.py"]


def _staged():
    return os.path.exists(os.path.join(REF, "SHA256SUMS"))


def test_vendored_reference_is_a_byte_copy():
    if not _staged():
        if os.path.exists("/root/reference/ficp.py"):
            subprocess.check_call([os.path.join(REPO, "tools", "vendor_ref.sh")])
        else:
            pytest.skip("oracle/_ref not staged and /root/reference absent (run tools/vendor_ref.sh in the build container)")
    sums = dict(line.split()[::-1] for line in open(os.path.join(REF, "SHA256SUMS")))
    assert sorted(sums) == sorted(FILES)
    for f in FILES:
        data = open(os.path.join(REF, f), "rb").read()
        assert hashlib.sha256(data).hexdigest() == sums[f], f
        if os.path.exists(os.path.join("/root/reference", f)):
            assert data == open(os.path.join("/root/reference", f), "rb").read(), f"{f} differs from /root/reference"


def test_repo_has_no_copy_of_reference_sources_in_history():
    out = subprocess.run(["git", "-C", REPO, "ls-files", "oracle/_ref"], capture_output=True, text=True)
    if out.returncode == 0:
        assert out.stdout.strip() == "", "oracle/_ref must stay out of git history"


RUNNER = r"""
import sys
sys.path.insert(0, {repo!r})
import ficp, coregistrationgame_b200.ficp as ours
assert ficp.FractionalICP is ours.FractionalICP, ficp.__file__
import pytest
class Check:
    def pytest_collection_modifyitems(self, items):
        import ficp as f2
        assert f2.FractionalICP is ours.FractionalICP
        for it in items:
            mod = it.module
            if hasattr(mod, "FractionalICP"):
                assert mod.FractionalICP is ours.FractionalICP, "reference test bound the wrong FractionalICP"
sys.exit(pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir", {ref!r}, "-c", "/dev/null",
                      {t1!r}, {t2!r}], plugins=[Check()]))
"""


@pytest.mark.gpu
def test_reference_test_files_run_verbatim_against_the_b200_ficp():
    if not _staged():
        pytest.skip("oracle/_ref not staged (tools/vendor_ref.sh runs in the build container; see __graft_entry__.build)")
    code = RUNNER.format(repo=REPO, ref=REF, t1=os.path.join(REF, "tests", "test_ficp.py"),
                         t2=os.path.join(REF, "tests", "test_rigid_2d_operations.py"))
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=REF, env=env, timeout=600)
    tail = (res.stdout + res.stderr)[-3000:]
    assert res.returncode == 0, tail
    assert "8 passed" in res.stdout, tail      # 5 in test_ficp.py + 3 in test_rigid_2d_operations.py
