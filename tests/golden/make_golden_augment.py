"""Golden vectors for the augmentation schemes other than log1p_drop (scnym/dataprep.py:730-754), produced by running the
UNMODIFIED reference's own `AUGMENTATION_SCHEMES` objects in the build container while RECORDING every random draw they
make (nothing is replaced): `np.random.random` / `np.random.choice` (GeneMasking), `torch.rand` (PoissonSample depth),
`torch.distributions.Poisson.sample` (PoissonSample), `F.dropout` (InputDropout; the keep mask is recovered from its output).
    python tests/golden/make_golden_augment.py  ->  tests/golden/augment_schemes.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from oracle.scnym_oracle import synth_csr, csr_to_dense  # noqa: E402

ref = load_reference()
DP = ref.dataprep


class Recorder(object):
    def __init__(self):
        self.choice, self.random, self.rand, self.pois, self.drop_out = [], [], [], [], []

    def __enter__(self):
        self.o_choice, self.o_random = np.random.choice, np.random.random
        self.o_rand, self.o_sample, self.o_dropout = torch.rand, torch.distributions.Poisson.sample, torch.nn.functional.dropout
        rec = self

        def choice(*a, **k):
            out = rec.o_choice(*a, **k)
            rec.choice.append(np.array(out).copy())
            return out

        def random(*a, **k):
            out = rec.o_random(*a, **k)
            rec.random.append(out)
            return out

        def rand(*a, **k):
            out = rec.o_rand(*a, **k)
            rec.rand.append(out.clone())
            return out

        def sample(self_, *a, **k):
            out = rec.o_sample(self_, *a, **k)
            rec.pois.append(out.clone())
            return out

        def dropout(x, *a, **k):
            out = rec.o_dropout(x, *a, **k)
            rec.drop_out.append((x.clone(), out.clone()))
            return out
        np.random.choice, np.random.random = choice, random
        torch.rand, torch.distributions.Poisson.sample, torch.nn.functional.dropout = rand, sample, dropout
        return self

    def __exit__(self, *a):
        np.random.choice, np.random.random = self.o_choice, self.o_random
        torch.rand, torch.distributions.Poisson.sample, torch.nn.functional.dropout = self.o_rand, self.o_sample, self.o_dropout


def main():
    G, B = 240, 10
    ip, ii, dd, _ = synth_csr(B, G, 0.2, 3, seed=17)
    x = csr_to_dense(ip, ii, dd, G)
    out = {"x": x.numpy(), "meta_G": G, "meta_B": B}
    # raw counts for count_poisson
    rng = np.random.default_rng(5)
    counts = torch.from_numpy((rng.poisson(1.5, size=(B, G)) * (rng.random((B, G)) < 0.3)).astype(np.float32))
    out["counts"] = counts.numpy()
    torch.manual_seed(123)
    for name in ("log1p_mask", "log1p_poisson", "log1p_poisson_drop", "count_poisson"):
        scheme = DP.AUGMENTATION_SCHEMES[name]
        inp = counts if name == "count_poisson" else x
        tries = 0
        while True:
            np.random.seed(100 + tries)
            with Recorder() as rec:
                y = scheme({"input": inp.clone(), "output": torch.zeros(B)})["input"]
            if name != "log1p_mask" or len(rec.choice) > 0:
                break                      # log1p_mask: retry until the batch-level decision applies the mask
            out["log1p_mask_noop/out"] = y.numpy()     # the no-op branch is a golden too (keep all ones)
            tries += 1
        out[name + "/out"] = y.numpy()
        if name == "log1p_mask":
            keep = np.ones((B, G), dtype=np.uint8)
            assert len(rec.choice) == B
            for r, idx in enumerate(rec.choice):
                keep[r, idx] = 0
            out[name + "/keep"] = keep
            out[name + "/n_masked"] = np.array([len(c) for c in rec.choice])
        if "poisson" in name:
            assert len(rec.pois) == 1
            out[name + "/pois"] = rec.pois[0].numpy().astype(np.float32)
        if name == "log1p_poisson_drop":
            assert len(rec.rand) >= 1 and len(rec.drop_out) == 1
            out[name + "/row_u"] = rec.rand[0].numpy()
            xin, xout = rec.drop_out[0]
            keep = ((xout != 0) | (xin == 0)).numpy().astype(np.uint8)   # where the input is 0 the decision does not matter
            out[name + "/keep"] = keep
        print(name, "ok", y.shape, float(torch.nan_to_num(y).abs().max()))
    np.savez_compressed(os.path.join(HERE, "augment_schemes.npz"), **out)


if __name__ == "__main__":
    main()
