#!/usr/bin/env python
"""Per-kernel SASS summary of the engine object (no GPU needed): instruction count, top opcodes and every
memory / synchronisation mnemonic with its count.   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "rl_stdp_b200", "csrc", "_obj", "snn_f32.o")
MEM = re.compile(r"^(LD|ST|ATOM|RED|UBLKCP|SYNCS|MEMBAR|FENCE|BAR|CCTL|ERRBAR|UTMA|NANOSLEEP|MATCH|VOTE|SHFL|REDUX|LDGSTS|LDGDEPBAR|DEPBAR)")
WANT = ("chain_", "learn_pass_tma", "neuron_kernel", "neuron_remote", "synapse_kernel", "ltp_kernel", "synapse_wave", "ltp_wave",
        "remote_kernel", "learn_kernel")


def main():
    out = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    print("SASS of rl_stdp_b200/csrc/snn_f32.cu for sm_100a (cuobjdump -sass of the object built by rl_stdp_b200/build.py, nvcc 12.9):")
    print("per kernel the instruction count, the most frequent opcodes and every memory / synchronisation mnemonic with its count.")
    print("UBLKCP.* = cp.async.bulk (TMA bulk copy engine), SYNCS.* = mbarrier, LDGSTS = cp.async (global -> shared, no register),")
    print("REDG.* = red.global (fire-and-forget reductions), ATOMG.* = returning atomics (list positions, dispensers, log")
    print("appends), LDG.E.*.STRONG.GPU / .SYS = ld.acquire / ld.relaxed of the hand-offs, LDL / STL = local memory (spills).\n")
    name, ops = None, []
    def flush():
        if name and any(w in name for w in WANT):
            c = collections.Counter(o.split(".")[0] for o in ops)
            m = collections.Counter(o for o in ops if MEM.match(o))
            print("== " + name)
            print("   instructions: %d   top: %s" % (len(ops), ", ".join("%s %d" % kv for kv in c.most_common(12))))
            print("   memory / sync mnemonics: " + ", ".join("%s %d" % kv for kv in sorted(m.items())))
    for line in out.splitlines():
        mm = re.search(r"Function : (\S+)", line)
        if mm:
            flush(); name, ops = mm.group(1), []
            continue
        mm = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if mm:
            ops.append(mm.group(1))
    flush()


if __name__ == "__main__":
    main()
