"""CLI plumbing of the reference's `gin_wrap` (src/utils.py:58-80): `<script> save_path config
[--bindings ...]`, several configs joined by '#', bindings separated by '#'; stdout/stderr are
teed into `<save_path>/stdout.txt` / `stderr.txt`."""
from __future__ import annotations

import argparse
import os
import sys
from contextlib import contextmanager

from . import gin_lite


class _Tee:
    def __init__(self, *streams):
        self.streams = streams

    def write(self, data):
        for s in self.streams:
            s.write(data)

    def flush(self):
        for s in self.streams:
            s.flush()


@contextmanager
def tee_std_streams(stdout_path, stderr_path):
    with open(stdout_path, 'a', 1) as out, open(stderr_path, 'a', 1) as err:
        old = sys.stdout, sys.stderr
        sys.stdout, sys.stderr = _Tee(old[0], out), _Tee(old[1], err)
        try:
            yield
        finally:
            sys.stdout, sys.stderr = old


def gin_wrap(fnc, argv=None):
    ap = argparse.ArgumentParser(description=fnc.__doc__)
    ap.add_argument("save_path")
    ap.add_argument("config", help="gin file(s), '#'-separated")
    ap.add_argument("-b", "--bindings", default="", help="'#'-separated Name.param=value overrides")
    args = ap.parse_args(argv)
    gin_lite.parse_config_files_and_bindings(args.config.split("#"), args.bindings.replace("#", "\n"))
    os.makedirs(args.save_path, exist_ok=True)
    with tee_std_streams(os.path.join(args.save_path, "stdout.txt"), os.path.join(args.save_path, "stderr.txt")):
        return fnc(args.save_path)
