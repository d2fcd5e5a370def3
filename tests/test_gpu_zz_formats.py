"""The whole command on the device with every output file: usearch12_b200_cli (FASTA in, kernels, HitMgr
mirror, OutputSink / DBHitSink) against the files the unmodified reference binary wrote for the same command
line (tests/golden/fmt_*, tools/make_golden_formats.py).  The formats alone are pinned without a GPU by
tests/test_formats_cpu.py; here the hits come from the CUDA path and the database letters from the index."""
import os
import subprocess
import sys

import pytest

from tests import util
from tests.test_formats_cpu import OUT_FLAGS, check_outputs

sys.path.insert(0, os.path.join(util.ROOT, "tools"))
import make_golden_formats as M  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(M.VARIANTS))
def test_cli_writes_the_reference_files(name, tmp_path):
    from usearch12_b200 import build
    cli = build.build_cli()
    tmp = str(tmp_path)
    q, d = M.write_inputs(name, tmp)
    cmd_name, _, _, _, opts, fields = M.VARIANTS[name]
    paths = {k: os.path.join(tmp, "o." + k) for k in OUT_FLAGS}
    cmd = [cli, "-" + cmd_name, q, "-db", d, "-quiet", "-userfields", fields] + opts
    for k, flag in OUT_FLAGS.items():
        cmd += [flag, paths[k]]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    check_outputs(name, paths)
