"""SASS fingerprint of one kernel inside a built library: the stamp that ties an ncu measurement to the code it measured.

bench.py reports `roofline.traffic` (DRAM bytes per launch from an `ncu --set full` capture) only while the kernel's SASS
still hashes to the value recorded with the capture; after any change to the kernel the stale number is dropped instead
of being repeated.  `python scripts/sass_hash.py <lib.so> <regex on the mangled name>` prints the hash.
"""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess
import sys


def _cuobjdump():
    for cand in (shutil.which("cuobjdump"), "/usr/local/cuda/bin/cuobjdump"):
        if cand and os.path.exists(cand):
            return cand
    return None


def kernel_sass_hash(lib_path: str, name_regex: str):
    """sha256 over the instruction text (addresses and encodings stripped) of every function whose mangled name matches."""
    tool = _cuobjdump()
    if not tool or not os.path.exists(lib_path):
        return None
    try:
        out = subprocess.run([tool, "-sass", lib_path], capture_output=True, text=True, timeout=120).stdout
    except Exception:  # noqa: BLE001
        return None
    pat = re.compile(name_regex)
    h = hashlib.sha256()
    take = False
    matched = 0
    for line in out.splitlines():
        s = line.strip()
        if s.startswith("Function :"):
            take = bool(pat.search(s))
            matched += take
            continue
        if not take or not s.startswith("/*"):
            continue
        m = re.match(r"/\*[0-9a-f]+\*/\s+(.*?);", s)       # "/*0590*/  INSN operands ;  /* encoding */"
        if m:
            h.update(m.group(1).strip().encode())
            h.update(b"\n")
    return h.hexdigest() if matched else None


if __name__ == "__main__":
    print(kernel_sass_hash(sys.argv[1], sys.argv[2]))
