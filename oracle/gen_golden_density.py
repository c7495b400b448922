"""Golden vectors for the density-tensor build: runs the REFERENCE's own general-XRCC/build_density_tensors.py (and its
density_tensors.c / H_contractions.c, compiled where they lie by the qode shim's import_C) in this container.

TEST INFRASTRUCTURE; needs /root/reference.     python oracle/gen_golden_density.py
Inputs: oracle.density_oracle.make_states (seeded): 4 spatial orbitals, 1 frozen core, charges 0/+1/-1 around 4 electrons
(15/6/20 configurations, 3/2/3 states), V = seeded standard normal (8^4)."""
import os
import sys
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REFERENCE = os.environ.get("XR_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(HERE, "qode_shim"))
sys.path.insert(0, os.path.join(REFERENCE, "general-XRCC"))

from oracle import density_oracle as do      # noqa: E402

CASE = dict(n_orbs=4, n_core=1, n_elec_ref=4, n_states={0: 3, +1: 2, -1: 3}, seed=7, v_seed=70)


def inputs():
    z_lists = do.make_states(CASE["n_orbs"], CASE["n_core"], CASE["n_elec_ref"], CASE["n_states"], CASE["seed"])
    n = 2 * CASE["n_orbs"]
    V = numpy.random.default_rng(CASE["v_seed"]).standard_normal((n, n, n, n))
    return z_lists, V


def main():
    import build_density_tensors as ref
    z_lists, V = inputs()
    densities, total = ref.build_density_tensors(z_lists, CASE["n_orbs"], V, CASE["n_core"])
    out = {"total_size": numpy.array(total)}
    for op, blocks in densities.items():
        for (bra, ket), nested in blocks.items():
            out["%s_%d_%d" % (op, bra, ket)] = numpy.array(nested)
    numpy.savez_compressed(os.path.join(REPO, "tests", "golden", "density_tensors.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
