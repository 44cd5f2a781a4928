#!/usr/bin/env python3
"""Extract the golden vectors of SURVEY.md §8(c) from the reference's legacy fixtures.

Reads  /root/reference/test/testin/input_pair_{3206,1003}.tsv  and  mcmc_{3206,1003}.tsv
(the command lines that produced them are in /root/reference/test/pred.jl:3,22) and writes
compact CSV fixtures under tests/golden/.  Only runs in the build container (the reference
tree does not exist on the GPU box); the outputs are committed.

Columns kept from the input table: PersonID, StoolPairs, nutrient (the X columns of the
kernel-program, Appendix D of SURVEY.md) and bug (the response).  The chain tables are
comma separated despite their extension; all columns are kept.
"""
import csv
import os
import sys

REF = "/root/reference/test/testin"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def convert(tag):
    with open(f"{REF}/input_pair_{tag}.tsv") as f:
        rows = list(csv.DictReader(f, delimiter="\t"))
    with open(f"{OUT}/input_pair_{tag}.csv", "w") as f:
        f.write("PersonID,StoolPairs,nutrient,bug\n")
        for r in rows:
            f.write(f"{r['PersonID']},{r['StoolPairs']},{r['nutrient']},{r['bug']}\n")
    with open(f"{REF}/mcmc_{tag}.tsv") as f:
        lines = f.read().strip().split("\n")
    names = {"θc[σ2]": "var1", "θc[σ2_2]": "var2", "θc[σ2_3]": "var3", "θc[σ2_4]": "var4",
             "θl[η]": "eta", "lπ": "lpi"}
    hdr = [names[h] for h in lines[0].split(",")]
    with open(f"{OUT}/mcmc_{tag}.csv", "w") as f:
        f.write(",".join(hdr) + "\n")
        for ln in lines[1:]:
            f.write(ln + "\n")
    print(tag, len(rows), "rows;", len(lines) - 1, "chain rows")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are already committed")
    for tag in ("3206", "1003"):
        convert(tag)
