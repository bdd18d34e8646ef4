#!/usr/bin/env python
"""Summarises `nvcc -Xptxas -v` logs: kernel, registers, shared memory, spills (tools/, not product)."""
import re
import subprocess
import sys


def main(paths):
    for path in paths:
        text = open(path).read()
        cur = None
        for line in text.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = {"name": m.group(1), "spill": "0/0"}
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                cur["spill"] = f"{m.group(2)}/{m.group(3)}"
            m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", line)
            if m and cur:
                name = cur["name"]
                name = re.sub(r"^__nv_static_\d+__\w+?__(_Z)", r"\1", name)
                dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
                dem = re.sub(r"acgpu::\(anonymous namespace\)::", "", dem)
                dem = re.sub(r"\(.*\)$", "", dem).replace("void ", "")
                print(f"{dem:60s} regs={m.group(1):>3s} smem={m.group(2) or 0:>6} spill={cur['spill']}")
                cur = None


if __name__ == "__main__":
    main(sys.argv[1:])
