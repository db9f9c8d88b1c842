"""Command line front end: the JSON wire format of ``bqa_cli`` (reference src/bqa/cli.py:17-176) on the B200 engine.

    python -m bqa_b200.cli [-i config.json] [-o result.json] [-l LOG_LEVEL] [--precision single|double] [--device cuda:0]
                           [--checkpoint state.npz --checkpoint-every N [--resume]]

Same contract as the reference tool: the config is read from ``-i`` (a ``*.json`` path relative to the current working
directory) or from stdin, ``run_qa`` is called on it, and the result list
``[["bloch_vectors", [[x, y, z], ...]] | ["measurement_outcomes", [+1 | -1, ...]], ...]`` is written as JSON to ``-o``
or to stdout; ``-l`` takes DEBUG / INFO / WARNING / ERROR (default INFO).  Any failure prints the chain of error
messages to stderr and exits with status 1 (reference cli.py:158-176); an unknown argument prints a hint and exits with
status 1 (:36-37, :88-90).  ``--precision`` / ``--device`` / ``--checkpoint*`` / ``--resume`` are additions of this engine (the reference selects the
precision with the BQA_PRECISION environment variable, utils.py:9-20, which is honoured here too)."""
from __future__ import annotations

import json
import logging
import os
import sys
from pathlib import Path

LOG_LEVELS = ("DEBUG", "INFO", "WARNING", "ERROR")
DEFAULT_LOG_LEVEL = "INFO"
PRECISIONS = ("single", "double")


class CliError(Exception):
    """A problem with the command line, the input file or the output file (not with the computation)."""


def usage() -> str:
    cwd = os.getcwd()
    return (
        "usage: python -m bqa_b200.cli [options]\n\n"
        "options:\n"
        f"  -i | --input PATH            *.json config, relative to {cwd}; stdin when absent\n"
        f"  -o | --output PATH           *.json file for the results, relative to {cwd}; stdout when absent\n"
        f"  -l | --log-level LEVEL       one of {', '.join(LOG_LEVELS)} (default {DEFAULT_LOG_LEVEL})\n"
        f"       --precision PRECISION   one of {', '.join(PRECISIONS)} (default: BQA_PRECISION, else single)\n"
        "       --device DEVICE         CUDA device of the engine (default cuda:0)\n"
        f"       --checkpoint PATH       *.npz file, relative to {cwd}: state + schedule position, written every\n"
        "                               --checkpoint-every N instructions (default 0: never)\n"
        "       --resume                continue from --checkpoint if the file exists\n"
        "  -h | --help                  show this message and exit")


def _json_path(text: str) -> Path:
    try:
        path = (Path(os.getcwd()) / Path(text)).resolve()
    except (RuntimeError, OSError) as e:
        raise CliError(f"cannot resolve the path {text}") from e
    if path.suffix != ".json":
        raise CliError(f"{path} must have the .json suffix")
    return path


def parse_args(argv: list[str]) -> dict:
    """Options as a dict: input / output (a Path, or None for stdin / stdout), log-level, precision, device.
    ``-h`` prints the usage and exits 0; an unknown argument prints a hint and exits 1."""
    opts = {"input": None, "output": None, "log-level": DEFAULT_LOG_LEVEL, "precision": None, "device": None,
            "checkpoint": None, "checkpoint-every": 0, "resume": False}
    takes_value = {"-i": "input", "--input": "input", "-o": "output", "--output": "output", "-l": "log-level",
                   "--log-level": "log-level", "--precision": "precision", "--device": "device",
                   "--checkpoint": "checkpoint", "--checkpoint-every": "checkpoint-every"}
    args = iter(argv[1:])
    for arg in args:
        if arg in ("-h", "--help"):
            print(usage())
            sys.exit(0)
        if arg == "--resume":
            opts["resume"] = True
            continue
        if arg not in takes_value:
            print(f"Invalid command line argument {arg}, run `{argv[0]} --help` to get the documentation")
            sys.exit(1)
        key = takes_value[arg]
        value = next(args, None)
        if value is None:
            raise CliError(f"no value after the {arg} key")
        if key in ("input", "output"):
            opts[key] = _json_path(value)
        elif key == "log-level":
            if value not in LOG_LEVELS:
                raise CliError(f"{value} is not a logging level, must be one of {', '.join(LOG_LEVELS)}")
            opts[key] = value
        elif key == "checkpoint":
            try:
                path = (Path(os.getcwd()) / Path(value)).resolve()
            except (RuntimeError, OSError) as e:
                raise CliError(f"cannot resolve the path {value}") from e
            if path.suffix != ".npz":
                raise CliError(f"{path} must have the .npz suffix")
            opts[key] = str(path)
        elif key == "checkpoint-every":
            if not value.isdigit():
                raise CliError(f"{value} is not a number of instructions")
            opts[key] = int(value)
        elif key == "precision":
            if value not in PRECISIONS:
                raise CliError(f"{value} is not a precision, must be one of {', '.join(PRECISIONS)}")
            opts[key] = value
        else:
            opts[key] = value
    return opts


def read_config(src: Path | None):
    try:
        if src is None:
            return json.load(sys.stdin)
        with src.open("r") as f:
            return json.load(f)
    except (json.JSONDecodeError, UnicodeError, OSError) as e:
        raise CliError("Error while parsing json data") from e


def write_result(dst: Path | None, result) -> None:
    try:
        if dst is None:
            json.dump(result, sys.stdout)
            return
        with dst.open("w") as f:
            json.dump(result, f)
    except OSError as e:
        raise CliError("Error while writing result") from e


def format_error(e: BaseException | None) -> str:
    msgs = []
    while e is not None:
        msgs.append(str(e))
        e = e.__cause__
    return "\ncaused by: ".join(msgs)


def main(argv: list[str] | None = None, run=None) -> int:
    """Returns the exit status.  ``run``: the function called on the config (tests inject one; default ``run_qa``)."""
    argv = list(sys.argv if argv is None else argv)
    try:
        opts = parse_args(argv)
        logging.basicConfig(level=getattr(logging, opts["log-level"]), format="%(asctime)s [%(levelname)s] %(message)s",
                            datefmt="%Y-%m-%d %H:%M:%S")
        config = read_config(opts["input"])
        if run is None:
            from .core import run_qa

            def run(cfg):
                return run_qa(cfg, precision=opts["precision"], device=opts["device"], checkpoint=opts["checkpoint"],
                              checkpoint_every=opts["checkpoint-every"], resume=opts["resume"])
        write_result(opts["output"], run(config))
    except Exception as e:                      # like the reference: every failure is reported, status 1
        print(format_error(e), file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
