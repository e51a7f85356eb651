#!/usr/bin/env python
"""SASS evidence of the B200-native kernels: disassembles libcgb200.so (cuobjdump -sass) and writes, per
kernel, the counts of the mnemonics that matter plus short excerpts around them.
    python profiles/sass_excerpt.py > profiles/r02/sass_excerpt.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "conjugate-gradient_b200", "libcgb200.so")
KERNELS = [
    ("_ZN3cgb17cg_persist_kernelILi8ELi2ELi512ELi3ELi2EEEvNS_11PersistArgsE",
     "cg_persist_kernel<8,2,512,3> -- the whole CG loop, one cooperative launch (csrc/persist.cu)"),
    ("_ZN3cgb15gemv_tma_kernelILi8ELi2ELi512ELi3ELi1ELi0EEEvNS_8GemvArgsE",
     "gemv_tma_kernel<8,2,512,3> -- mat-vec of the graph schedule, init and DEBUG mat-vecs (csrc/gemv.cu)"),
    ("_ZN3cgb21gemv_tensormap_kernelILi8ELi1ELi1024ELi3EEEvNS_8GemvArgsE14CUtensorMap_st",
     "gemv_tensormap_kernel<8,1,1024,3> -- the A/B arm only: 2-D tensor-map producer (profiles/r02/tensormap_ab.md)"),
]
WHAT = [
    ("UBLKCP", "cp.async.bulk.shared::cluster.global.mbarrier (1-D TMA bulk copy of a tile row / p slice)"),
    ("UBLKPF", "cp.async.bulk.prefetch.L2 (A tiles of the next mat-vec into L2)"),
    ("UTMALDG", "cp.async.bulk.tensor.2d (tensor-map TMA) -- only in the tm2d_* A/B variants; the product copies tile rows as 1-D bulk rows"),
    ("SYNCS", "mbarrier init / arrive / expect_tx / try_wait"),
    ("LDS.128", "128-bit shared-memory loads of A and p (conflict-free: lane l reads chunk l, l+32, ...)"),
    ("DFMA", "fp64 FMA -- CUDA cores; no tensor-core instruction (HMMA/UTC*MMA) in an HBM-bound GEMV"),
    ("STG.E.128.STRONG.SYS", "self-flagging 16-byte LL store {lo, tag, hi, tag} to a (peer-mapped) gather buffer"),
    ("LDG.E.128.STRONG.SYS", "poll of an LL entry (ld.volatile.v4)"),
    ("REDG.E.ADD", "red.release.gpu: a CTA's p chunks are published"),
    ("LDG.E.STRONG.GPU", "ld.acquire.gpu: the producer warp polls the p counter"),
    ("FENCE.VIEW.ASYNC", "fence.proxy.async.global: generic-proxy stores to p become visible to the TMA loads"),
    ("NANOSLEEP", "back-off inside the waits"),
    ("BAR.SYNC", "named barrier among the consumer warps (the producer warp never joins)"),
    ("SHFL", "warp butterflies (row dots, chunk256 trees)"),
    ("HMMA", "legacy tensor-core path -- must be 0"),
    ("UTC", "tcgen05 -- must be 0 (fp64, 0.25 flop/byte)"),
]


def main():
    print("# SASS evidence (cuobjdump -sass of conjugate-gradient_b200/libcgb200.so, sm_100a)\n")
    for sym, title in KERNELS:
        r = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True)
        lines = [ln.rstrip() for ln in r.stdout.splitlines() if re.search(r"/\*[0-9a-f]{4,6}\*/", ln)]
        text = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).strip() for ln in lines]
        print("## %s\n\n`%s`, %d SASS instructions\n" % (title, sym, len(text)))
        print("| mnemonic | count | what it is |\n|---|---|---|")
        for key, what in WHAT:
            n = sum(1 for t in text if key in t)
            print("| `%s` | %d | %s |" % (key, n, what))
        print()
        for key in ("UBLKCP", "UTMALDG", "UBLKPF", "STG.E.128.STRONG.SYS", "LDG.E.128.STRONG.SYS", "REDG.E.ADD", "FENCE.VIEW.ASYNC"):
            hits = [i for i, t in enumerate(text) if key in t]
            if not hits:
                continue
            i = hits[0]
            print("`%s` (first of %d):\n```" % (key, len(hits)))
            for t in text[max(0, i - 3):i + 3]:
                print(t)
            print("```")
        # the inner loop: the densest window of LDS.128 + DFMA
        best, bi = -1, 0
        for i in range(0, max(1, len(text) - 40)):
            w = text[i:i + 40]
            sc = sum(1 for t in w if "DFMA" in t) + sum(1 for t in w if "LDS.128" in t)
            if sc > best:
                best, bi = sc, i
        print("consumer inner loop (densest 40-instruction window: %d of LDS.128 / DFMA):\n```" % best)
        for t in text[bi:bi + 40]:
            print(t)
        print("```\n")


if __name__ == "__main__":
    main()
