"""Adds the -tabbedout fixtures (State2::OutputTab2, outputtab2.cpp) to tests/golden/: the UNMODIFIED reference binary run
with -threads 1 on the committed paired-end FASTQs.

    python tests/golden/make_golden_tab.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_py as O


def main():
    ufi = os.path.join(HERE, "ref.ufi")
    for out, extra in (("pe.tab", []), ("pe_veryfast.tab", ["-veryfast"])):
        O.run_reference(["-map2", os.path.join(HERE, "pe_1.fq"), "-reverse", os.path.join(HERE, "pe_2.fq"), "-ufi", ufi,
                         "-samout", "/tmp/golden_tab_unused.sam", "-tabbedout", os.path.join(HERE, out), "-threads", "1"] + extra)
        lines = open(os.path.join(HERE, out), "rb").read().split(b"\n")[:-1]
        print(out, len(lines), "pairs,", sum(1 for l in lines if l.split(b"\t")[3] != b"*"), "with a second pair")


if __name__ == "__main__":
    main()
