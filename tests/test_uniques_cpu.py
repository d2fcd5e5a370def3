"""-fastx_uniques argument checks that need no device (derepfull.cpp:214-218); the dereplication itself runs
on the device and is tested in tests/test_gpu_uniques.py."""
import subprocess

from usearch12_b200 import build


def test_fastx_uniques_refuses_output_option(tmp_path):
    cli = build.build_cli()
    r = subprocess.run([cli, "-fastx_uniques", "x.fa", "-output", "y.fa"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "Use -fastaout, not -output" in r.stdout


def test_fastx_uniques_fails_loudly_without_a_device(tmp_path):
    """No CPU path: without a CUDA device the command must stop with an error, not fall back."""
    from usearch12_b200 import capi
    if capi.lib().usb_device_count() > 0:
        import pytest
        pytest.skip("a CUDA device is present")
    src = tmp_path / "in.fa"
    src.write_text(">a\nACGT\n>b\nacgt\n")
    r = subprocess.run([build.build_cli(), "-fastx_uniques", str(src), "-fastaout", str(tmp_path / "o.fa")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode != 0 and "CUDA" in r.stdout
