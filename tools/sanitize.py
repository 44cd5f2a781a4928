#!/usr/bin/env python3
"""compute-sanitizer over one small call of every kernel.

    python tools/sanitize.py memcheck|racecheck|synccheck|initcheck [--big worker]

ONE tool per GPU-box call (profiling recipe).  Writes the tool's log to gpurun_out/sanitize_<tool>.log and a short
summary (error counts) to gpurun_out/sanitize_<tool>_summary.txt; the summaries are committed under profiles/."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tool = sys.argv[1]
    extra = sys.argv[2:]
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    log = os.path.join(out_dir, f"sanitize_{tool}.log")
    # plain run first: the tool only ever sees a program that has just exited 0 on this box
    plain = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_target.py"), *extra], capture_output=True, text=True)
    if plain.returncode != 0:
        print("plain run failed; not running the sanitizer\n", plain.stdout[-2000:], plain.stderr[-2000:])
        return 1
    cmd = ["compute-sanitizer", "--tool", tool, "--log-file", log, "--error-exitcode", "0", "--launch-timeout", "0",
           sys.executable, os.path.join(ROOT, "tools", "sanitize_target.py"), *extra]
    if tool == "memcheck":
        cmd[3:3] = ["--leak-check", "no"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    text = open(log).read() if os.path.exists(log) else ""
    m = re.findall(r"ERROR SUMMARY: (\d+) error", text)
    hazards = len(re.findall(r"(Race reported|hazard|Invalid __|Barrier error|Uninitialized)", text))
    summary = (f"compute-sanitizer --tool {tool} {' '.join(extra)} on tools/sanitize_target.py\n"
               f"target exit code {r.returncode}; target said: {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else '(nothing)'}\n"
               f"ERROR SUMMARY: {m[-1] if m else '?'} errors; {hazards} hazard / invalid-access records in the log\n")
    lines = [ln for ln in text.splitlines() if "=========" in ln][:60]
    summary += "\n".join(lines[:40]) + "\n"
    open(os.path.join(out_dir, f"sanitize_{tool}_summary.txt"), "w").write(summary)
    print(summary)
    return 0


if __name__ == "__main__":
    sys.exit(main())
