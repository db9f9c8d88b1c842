"""Command line front end (bqa_b200/cli.py) against the contract of the reference tool (src/bqa/cli.py): options,
JSON in / JSON out, stdin / stdout defaults, error reporting with exit status 1."""
import io
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import instances
from bqa_b200 import cli

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def json_config(cfg: dict) -> dict:
    """The list form the JSON wire format uses (reference examples/cli_examples/small_ibm_heavy_hex.json)."""
    out = dict(cfg)
    out["nodes"] = [[int(n), float(h)] for n, h in cfg["nodes"].items()]
    out["edges"] = [[[int(a), int(b)], float(j)] for (a, b), j in cfg["edges"].items()]
    return out


def test_parse_args_defaults_and_paths(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    opts = cli.parse_args(["prog"])
    assert opts == {"input": None, "output": None, "log-level": "INFO", "precision": None, "device": None,
                    "checkpoint": None, "checkpoint-every": 0, "resume": False}
    opts = cli.parse_args(["prog", "--checkpoint", "run.npz", "--checkpoint-every", "50", "--resume"])
    assert opts["checkpoint"] == str((tmp_path / "run.npz").resolve()) and opts["checkpoint-every"] == 50 and opts["resume"]
    opts = cli.parse_args(["prog", "-i", "a.json", "--output", "sub/b.json", "-l", "DEBUG", "--precision", "double",
                           "--device", "cuda:1"])
    assert opts["input"] == (tmp_path / "a.json").resolve() and opts["output"] == (tmp_path / "sub" / "b.json").resolve()
    assert (opts["log-level"], opts["precision"], opts["device"]) == ("DEBUG", "double", "cuda:1")


@pytest.mark.parametrize("argv,msg", [
    (["prog", "-i", "config.yaml"], ".json suffix"),
    (["prog", "-l", "LOUD"], "logging level"),
    (["prog", "--precision", "half"], "precision"),
    (["prog", "-o"], "no value"),
    (["prog", "--checkpoint", "state.bin"], ".npz suffix"),
    (["prog", "--checkpoint-every", "often"], "number of instructions"),
])
def test_bad_options_are_reported_with_status_1(argv, msg, capsys):
    assert cli.main(argv, run=lambda cfg: []) == 1
    assert msg in capsys.readouterr().err


def test_help_and_unknown_argument_exit_codes(capsys):
    with pytest.raises(SystemExit) as e:
        cli.parse_args(["prog", "--help"])
    assert e.value.code == 0 and "--input" in capsys.readouterr().out
    with pytest.raises(SystemExit) as e:
        cli.parse_args(["prog", "--frobnicate"])
    assert e.value.code == 1 and "--frobnicate" in capsys.readouterr().out


def test_json_in_json_out_through_files_and_streams(tmp_path, monkeypatch, capsys):
    monkeypatch.chdir(tmp_path)
    cfg = json_config(instances.cfg_ring24())
    (tmp_path / "cfg.json").write_text(json.dumps(cfg))
    seen = {}

    def run(c):
        seen["cfg"] = c
        return [["bloch_vectors", [[0.0, 0.5, -1.0]]], ["measurement_outcomes", [1, -1]]]
    assert cli.main(["prog", "-i", "cfg.json", "-o", "out.json", "-l", "ERROR"], run=run) == 0
    assert seen["cfg"] == cfg
    assert json.loads((tmp_path / "out.json").read_text()) == [["bloch_vectors", [[0.0, 0.5, -1.0]]],
                                                                ["measurement_outcomes", [1, -1]]]
    monkeypatch.setattr(sys, "stdin", io.StringIO(json.dumps(cfg)))          # stdin -> stdout
    assert cli.main(["prog", "-l", "ERROR"], run=run) == 0
    assert json.loads(capsys.readouterr().out) == [["bloch_vectors", [[0.0, 0.5, -1.0]]], ["measurement_outcomes", [1, -1]]]


def test_failures_print_the_cause_chain(tmp_path, monkeypatch, capsys):
    monkeypatch.chdir(tmp_path)
    (tmp_path / "broken.json").write_text("{ not json")
    assert cli.main(["prog", "-i", "broken.json"], run=lambda c: []) == 1
    err = capsys.readouterr().err
    assert "Error while parsing json data" in err and "caused by:" in err
    assert cli.main(["prog", "-i", "missing.json"], run=lambda c: []) == 1

    def run(c):
        try:
            raise KeyError("inner")
        except KeyError as e:
            raise RuntimeError("outer") from e
    (tmp_path / "ok.json").write_text("{}")
    assert cli.main(["prog", "-i", "ok.json"], run=run) == 1
    err = capsys.readouterr().err
    assert "outer" in err and "caused by: 'inner'" in err


def test_unknown_backend_is_rejected_not_silently_replaced(tmp_path, monkeypatch, capsys):
    """An unknown backend name fails loudly; the reference's own names are accepted with a warning (see below)."""
    monkeypatch.chdir(tmp_path)
    cfg = json_config(instances.cfg_ring24())
    cfg["backend"] = "tpu"
    (tmp_path / "cfg.json").write_text(json.dumps(cfg))
    assert cli.main(["prog", "-i", "cfg.json", "-l", "ERROR"]) == 1
    assert "backend" in capsys.readouterr().err


def test_reference_backend_names_compile_with_a_warning(caplog):
    """Every reference example / benchmark config sets "backend": "numpy" or "cupy" explicitly: they compile
    unmodified, are executed on b200, and say so (there is no CPU path to fall back to)."""
    import logging
    from bqa_b200.config import config_to_context
    for name in ("numpy", "cupy"):
        cfg = dict(instances.cfg_ring24(), backend=name)
        with caplog.at_level(logging.WARNING, logger="bqa_b200.config"):
            ctx = config_to_context(cfg)
        assert ctx.backend == "b200"
        assert any("b200" in r.message and name in r.message for r in caplog.records)
        caplog.clear()


@pytest.mark.gpu
def test_cli_end_to_end_matches_oracle(tmp_path):
    from oracle import bqa_oracle as O
    cfg = instances.cfg_ring24()
    (tmp_path / "cfg.json").write_text(json.dumps(json_config(cfg)))
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    p = subprocess.run([sys.executable, "-m", "bqa_b200.cli", "-i", "cfg.json", "-o", "out.json", "-l", "ERROR",
                        "--precision", "double"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    got = dict(json.loads((tmp_path / "out.json").read_text()))
    want = dict(O.run_qa(cfg))
    assert np.abs(np.array(got["bloch_vectors"]) - np.array(want["bloch_vectors"])).max() < 1e-8
    assert got["measurement_outcomes"] == want["measurement_outcomes"]
