"""Collects the reference's saved CartPole elites into one fixture (run in the build container, where
/root/reference exists; the fixture travels, the reference does not).

Source: Julia/Actors/cartpoleElites/sNESelites_cartpole_{1..17}.jld — `elites`, ten AbstractESIndividual
each (`genes` 160 x Float64, `fitness` 1 x Float64), written by the sNES runs of Actors/optim.jl and read
back by Actors/cartpole_tryout.jl:15.  Read with rl_stdp_b200/jld.py (no HDF5 library in the image).

        Julia/Actors/sNESelites_mountcar.jld — ten elites of the MountainCar-v0 runs (mountcar_fitness.jl),
        fitness -129 ... -117 (minus the decisions needed to reach the flag, ties counted twice).

Output: tests/golden/cartpole_elites.npz  genes [170,160] f64, fitness [170] f64, run [170] i32;
        tests/golden/mountcar_elites.npz  genes [10,160] f64, fitness [10] f64.
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from rl_stdp_b200 import jld  # noqa: E402

SRC = "/root/reference/Julia/Actors/cartpoleElites"


def main():
    genes, fit, run = [], [], []
    files = sorted(glob.glob(os.path.join(SRC, "sNESelites_cartpole_*.jld")), key=lambda p: int(p.split("_")[-1][:-4]))
    for f in files:
        k = int(f.split("_")[-1][:-4])
        for el in jld.load(f, "elites"):
            genes.append(np.asarray(el["genes"], np.float64))
            fit.append(float(np.ravel(el["fitness"])[0]))
            run.append(k)
    np.savez_compressed(os.path.join(HERE, "cartpole_elites.npz"), genes=np.array(genes), fitness=np.array(fit),
                        run=np.array(run, np.int32))
    print(len(genes), "elites from", len(files), "files; fitness values:", sorted(set(fit)))
    el = jld.load("/root/reference/Julia/Actors/sNESelites_mountcar.jld", "elites")
    np.savez_compressed(os.path.join(HERE, "mountcar_elites.npz"),
                        genes=np.array([np.asarray(e["genes"], np.float64) for e in el]),
                        fitness=np.array([float(np.ravel(e["fitness"])[0]) for e in el]))
    print(len(el), "mountain-car elites; fitness:", [float(np.ravel(e["fitness"])[0]) for e in el])


if __name__ == "__main__":
    main()
