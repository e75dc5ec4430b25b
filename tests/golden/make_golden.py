"""Provenance of the fixtures in tests/golden/.

dlsodes_example.json  transcribed from the reference's own printed output of the DLSODES
                      documentation example, /root/reference/src/opkdmain.f:2097-2133.
                      `python make_golden.py check` re-reads those lines from the reference
                      tree (this container only) and verifies the transcription.
network_golden.json   values of SURVEY.md App. B / App. E (computed while surveying by
                      following the Fortran loaders); the reference has no fixture for them.
inp/                  the reference's INPUT DATA files for the path (network tables and
                      initial abundances, /root/reference/inp/), copied verbatim because
                      /root/reference does not exist on the GPU box.  They are data, not source.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def check():
    ref = "/root/reference/src/opkdmain.f"
    lines = open(ref).read().splitlines()[2096:2133]
    nums = []
    for ln in lines:
        nums += [float(x) for x in re.findall(r"-?\d\.\d+e[+-]\d+", ln)]
    g = json.load(open(os.path.join(HERE, "dlsodes_example.json")))
    flat = []
    for o in g["outputs"]:
        flat += [o["t"], o["hu"]] + o["y"]
    assert len(nums) == len(flat), (len(nums), len(flat))
    assert all(abs(a - b) <= 1e-12 * max(1.0, abs(b)) for a, b in zip(nums, flat)), "transcription differs"
    for f in os.listdir(os.path.join(HERE, "inp")):
        a = open(os.path.join(HERE, "inp", f), "rb").read()
        b = open(os.path.join("/root/reference/inp", f), "rb").read()
        assert a == b, f
    print("golden fixtures match the reference tree")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "check":
        check()
    else:
        print(__doc__)
