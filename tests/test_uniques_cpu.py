"""-fastx_uniques (derepfull.cpp:130-236, derepresult.cpp:255-284,689-775):
host-only dereplication, byte-identical to the reference binary's -fastaout (tools/make_golden_uniques.py)."""
import gzip
import os
import subprocess

import pytest

from tests import util
from usearch12_b200 import build

VARIANTS = {
    "uniq_sizeout": ["-sizeout"],
    "uniq_relabel": ["-sizeout", "-relabel", "Uniq", "-minuniquesize", "2"],
    "uniq_plain": [],
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_fastx_uniques_byte_identical_to_reference(name, tmp_path):
    cli = build.build_cli()
    src = str(tmp_path / "in.fa")
    with gzip.open(os.path.join(util.GOLDEN, "uniq_in.fa.gz"), "rb") as f, open(src, "wb") as g:
        g.write(f.read())
    dst = str(tmp_path / "out.fa")
    r = subprocess.run([cli, "-fastx_uniques", src, "-fastaout", dst, "-quiet"] + VARIANTS[name], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    with gzip.open(os.path.join(util.GOLDEN, name + ".fa.gz"), "rb") as f:
        want = f.read()
    assert open(dst, "rb").read() == want


def test_fastx_uniques_refuses_output_option(tmp_path):
    cli = build.build_cli()
    r = subprocess.run([cli, "-fastx_uniques", "x.fa", "-output", "y.fa"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "Use -fastaout, not -output" in r.stdout
