#!/usr/bin/env python
"""Why a rank's share of the partitioned C3 batch runs below the full-batch rate: list-scheduling
simulation of the work-sorted shoot kernel.  Work per instance = Cash-Karp attempts + 150 x shots (from the CPU oracle on
the rank's block-cyclic shard); a warp takes as long as its slowest lane; S = resident warps of the GPU.  Prints the
perfectly balanced time, dynamic heavy-first hand-out at block sizes 32 / 64 / 128, and a static boustrophedon assignment.
CPU only.  usage: python tools/schedule_sim.py"""
import numpy as np, heapq, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import oracle as O
from cosmotransitions_b200 import potentials
N=1000000
c = 0.02 + 0.475 * (np.arange(N) + 0.5) / N
for G in (8,4):
    idx=np.nonzero((np.arange(N)//256)%G==0)[0]
    sub=idx[::4]   # quarter sample for speed, then replicate
    o=O.bounce_batch(O.make_config(alpha=3), potentials.quartic_family_params(c[sub]), 1.0, 0.0, nthreads=16)
    w=np.repeat(o["n_rkck"].astype(float)+150.0*o["n_shots"],4)[:len(idx)]   # work proxy: steps + per-shot overhead
    order=np.argsort(-w)
    ws=w[order]
    ng=(len(ws)+31)//32
    pad=np.zeros(ng*32); pad[:len(ws)]=ws
    gw=pad.reshape(ng,32).max(axis=1)     # warp time = slowest lane
    for S in (148*12, 148*16):
        # dynamic LPT at warp granularity (block=32) and at block=64/128 granularity
        def dyn(block_warps):
            nb=(ng+block_warps-1)//block_warps
            bw=np.zeros(nb)
            for b in range(nb): bw[b]=gw[b*block_warps:(b+1)*block_warps].max()
            slots=[0.0]*(S//block_warps); heapq.heapify(slots)
            end=0
            for b in range(nb):
                t=heapq.heappop(slots); t+=bw[b]; end=max(end,t); heapq.heappush(slots,t)
            return end
        # static snake over warps
        tot=np.zeros(S)
        for j in range((ng+S-1)//S):
            seg=gw[j*S:(j+1)*S]
            if j%2==0: tot[:len(seg)]+=seg
            else: tot[S-1-np.arange(len(seg))]+=seg
        ideal=gw.sum()/S
        print("G=%d S=%d groups=%d waves=%.2f ideal %.0f | dyn32 %.0f dyn64 %.0f dyn128 %.0f | snake %.0f | max single %.0f"%(G,S,ng,ng/S,ideal,dyn(1),dyn(2),dyn(4),tot.max(),gw.max()))
