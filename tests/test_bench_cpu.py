"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys the
driver reads, and ranks other than 0 of a multi-rank launch exit without work."""
import json
import os
import subprocess
import sys

from tests import util

BENCH = os.path.join(util.ROOT, "bench.py")
SMALL = ["--steps", "1", "--warmup", "0", "--ref-sample", "1500", "--reads", "20000", "--db", "1500"]


def test_reference_arm_line_has_the_contract_keys():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference"] + SMALL, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["metric"].startswith("query-seqs/sec usearch_global") and line["unit"] == "query-seqs/s"
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "query-seqs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2"] + SMALL, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
