"""Golden vectors for the gene re-indexing step in front of predict, from the UNMODIFIED reference
(`scnym.utils.build_classification_matrix`, /root/reference/scnym/utils.py:155-233).  Build container only:

    python tests/golden/make_golden_genes.py

Writes `tests/golden/gene_remap.npz`: the sample matrix (CSR), both gene lists and the reference's output for CSR
input and for dense input (as dense arrays)."""
import os
import sys

import numpy as np
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from oracle.scnym_oracle import synth_csr  # noqa: E402

ref = load_reference(modules=("utils",))


def case(name, n_cells, n_sample, n_model, n_common, density, seed, out):
    rng = np.random.default_rng(seed)
    indptr, indices, data, _ = synth_csr(n_cells, n_sample, density, 3, seed)
    X = sparse.csr_matrix((data, indices, indptr), shape=(n_cells, n_sample))
    sample_genes = np.array([f"s{seed}_{i}" for i in range(n_sample)], dtype=object)
    common = rng.choice(n_sample, size=n_common, replace=False)
    model_genes = np.array([f"m{seed}_{i}" for i in range(n_model)], dtype=object)
    slots = rng.choice(n_model, size=n_common, replace=False)     # where the common genes sit in the model's order
    model_genes[slots] = sample_genes[common]
    N_csr = ref.utils.build_classification_matrix(X, model_genes.astype(str), sample_genes.astype(str), gene_batch_size=7)
    N_dense = ref.utils.build_classification_matrix(X.toarray(), model_genes.astype(str), sample_genes.astype(str))
    assert sparse.isspmatrix_csr(N_csr) and isinstance(N_dense, np.ndarray)
    out.update({f"{name}_indptr": indptr, f"{name}_indices": indices, f"{name}_data": data,
                f"{name}_sample_genes": sample_genes.astype(str), f"{name}_model_genes": model_genes.astype(str),
                f"{name}_N_from_csr": N_csr.toarray().astype(np.float32), f"{name}_N_from_dense": N_dense.astype(np.float32)})


if __name__ == "__main__":
    out = {}
    case("subset", n_cells=37, n_sample=90, n_model=64, n_common=41, density=0.2, seed=11, out=out)       # model smaller
    case("superset", n_cells=23, n_sample=50, n_model=130, n_common=50, density=0.3, seed=12, out=out)    # all sample genes kept
    case("disjoint", n_cells=9, n_sample=40, n_model=33, n_common=0, density=0.25, seed=13, out=out)      # nothing survives
    np.savez_compressed(os.path.join(HERE, "gene_remap.npz"), **out)
    print("wrote gene_remap.npz:", {k: v.shape for k, v in out.items() if k.endswith("N_from_csr")})
