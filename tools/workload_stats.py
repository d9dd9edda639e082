#!/usr/bin/env python
"""Workload statistics of the K2 benchmark (Kuka vs scene16) computed with the CPU
oracle: survival per link, cull selectivity, SAT outcomes.  Design aid only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import acrobotics_b200 as ab
from acrobotics_b200.synthetic import scene16, random_configs
from oracle import np_oracle as no

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
spec = no.spec_from_robot(ab.Kuka())
sh, sR, sT = no.scene_arrays(scene16())
q = random_configs(N, 6, seed=3)
Rw, cw = no.world_boxes(spec, q)          # (N,6,3,3), (N,6,3)
halfs = np.array([b.half for b in spec.boxes])
rad = np.linalg.norm(halfs, axis=1)
srad = np.linalg.norm(sh, axis=1)
# pairwise SAT
hit, g = no.sat(halfs[None, :, None, :], Rw[:, :, None], cw[:, :, None], sh[None, None], sR[None, None], sT[None, None])
# cull A: scene-box slabs vs link sphere
d = cw[:, :, None, :] - sT[None, None]                     # (N,6,S,3)
y = np.einsum("sij,nbsi->nbsj", sR, d)                      # box-frame coords
cullA = (np.abs(y) > sh[None, None] + rad[None, :, None, None]).any(-1)
# cull B: link-box slabs vs scene sphere
z = np.einsum("nbij,nbsi->nbsj", Rw, -d)
cullB = (np.abs(z) > halfs[None, :, None, :] + srad[None, None, :, None]).any(-1)
# sphere-sphere
dist = np.linalg.norm(d, axis=-1)
cullS = dist > rad[None, :, None] + srad[None, None]
# exact sphere(link) vs OBB(scene) distance
ex = np.maximum(np.abs(y) - sh[None, None], 0)
cullD = np.linalg.norm(ex, axis=-1) > rad[None, :, None]
link_hit = hit.any(-1)                                     # (N,6)
alive = np.ones(N, bool)
print("N", N, "scene hit rate", link_hit.any(1).mean())
tot = dict(A=0, AB=0, D=0, DB=0, S=0, hits=0, links=0)
for b in range(6):
    a = alive
    na = a.sum()
    cA = (~cullA[a, b]).sum(1); cAB = (~(cullA | cullB)[a, b]).sum(1); cD = (~cullD[a, b]).sum(1)
    cDB = (~(cullD | cullB)[a, b]).sum(1); cS = (~cullS[a, b]).sum(1)
    print(f"link {b}: alive {na/N:.3f}  cand/alive A {cA.mean():.2f} (max {cA.max()}) AB {cAB.mean():.2f} D {cD.mean():.2f} DB {cDB.mean():.2f} sph {cS.mean():.2f} "
          f"P(any cand A) {(cA>0).mean():.2f} DB {(cDB>0).mean():.2f}  true-hit pairs/alive {hit[a,b].sum(1).mean():.2f}  P(link hits) {link_hit[a,b].mean():.3f}")
    tot["links"] += na; tot["A"] += cA.sum(); tot["AB"] += cAB.sum(); tot["D"] += cD.sum(); tot["DB"] += cDB.sum(); tot["S"] += cS.sum()
    alive = alive & ~link_hit[:, b]
print("alive after scene", alive.mean())
self_hit = no.is_in_self_collision(spec, q)
print("self hit | survived scene", self_hit[alive].mean(), " overall collision", (self_hit | ~alive).mean())
print({k: v / N for k, v in tot.items()})
