"""Soak of the probe kernel's first look under the lock-step emulator (tests/emu) against the oracle: three genomes (plain,
repeat-rich with tandem arrays and segmental duplications, soft-masked with an N run) x six kinds of read pairs (60 - 250 bases,
0 - 5 % substitutions, Search4 / Search5, damaged / N-containing / lower-case mates).  Test infrastructure, CPU only:
    python tools/first_look_soak.py        -> one line per case, 0 mismatching fields expected (7 200 pairs, about a minute)"""
import os, sys, numpy as np, time, subprocess, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
from oracle import oracle_py as oracle
from urmap_b200 import synth
import emu_py
import tempfile
W = tempfile.mkdtemp(prefix='urmb_soak_')
tot=0; bad=0
for gi,(glen, rf, tandem, segdup, nrun) in enumerate([(800_000,0.10,0,0,[]),(600_000,0.30,15,4,[(0,0.3,500)]),(500_000,0.05,30,2,[(1,0.5,3000)])]):
    g = synth.make_genome(glen, n_contigs=3, seed=100+gi, repeat_frac=rf, tandem=tandem, segdup=segdup, n_runs=nrun, lower_frac=0.05 if gi==2 else 0.0)
    fa, ufi = f"{W}/g{gi}.fa", f"{W}/g{gi}.ufi"
    g.write_fasta(fa)
    r = subprocess.run([os.path.join(ROOT, 'urmap_b200', 'bin', 'urmap_b200'),'-make_ufi',fa,'-output',ufi,'-quiet'],capture_output=True,text=True); assert r.returncode==0, r.stderr
    ix = oracle.Index(ufi)
    for ci,(rl, sub, indel, pm) in enumerate([(150,0.01,0.001,4),(150,0.003,0.0,5),(250,0.01,0.001,4),(100,0.02,0.002,4),(150,0.05,0.01,4),(60,0.0,0.0,4)]):
        r1, r2, names = synth.sim_pe(g, 400, rl, sub, indel, seed=7*gi+ci, ins_mean=max(300, 2*rl), ins_lo=max(200, rl+50))
        if ci % 2 == 0:   # a few damaged / N-containing / lower-case mates
            r2 = r2.copy(); r2[::37, :rl//2] = r1[::37, :rl//2]; r1 = r1.copy(); r1[::41, rl//3] = ord('N'); r1[5::53] = np.where(r1[5::53] < 91, r1[5::53] + 32, r1[5::53])
        f1,f2=f"{W}/p1.fq",f"{W}/p2.fq"
        synth.write_fastq(f1, r1, names, b"/1"); synth.write_fastq(f2, r2, names, b"/2")
        b1, b2 = oracle.ReadBatch.from_fastq(f1), oracle.ReadBatch.from_fastq(f2)
        br = 4 if pm==5 else -1
        o1, o2, uo = oracle.map_pe(ix, b1, b2, pe_method=pm, band_radius=br)
        seqs = np.concatenate([b1.seqs, b2.seqs]); offs = np.concatenate([b1.offs, b2.offs[1:] + b1.offs[-1]]).astype(np.uint32)
        re_, ue, cnt = emu_py.emu_map(ix, oracle.RESULT_DTYPE, seqs, offs, b1.n, True, pe_method=pm, band_radius=br)
        ro = np.concatenate([o1,o2]); nb=0
        for f in ("db_pos", "score", "best", "second", "mapq", "flags", "hit_count", "hsp_count"):
            nb += int((ro[f] != re_[f]).sum())
        po=[tuple(uo[r["path_off"]:r["path_off"]+r["path_runs"]].tolist()) for r in ro]; pe=[tuple(ue[r["path_off"]:r["path_off"]+r["path_runs"]].tolist()) for r in re_]
        nb += sum(1 for a,b in zip(po,pe) if a!=b)
        tot += b1.n; bad += nb
        print(f"genome {gi} cfg {ci} (rl {rl}, sub {sub}, pm {pm}): first look {emu_py.emu_first_look()}/{b1.n}, overflow {cnt[1]}, mismatches {nb}", flush=True)
    ix.close()
print("pairs", tot, "mismatching fields", bad)
import shutil
shutil.rmtree(W, ignore_errors=True)
sys.exit(1 if bad else 0)
