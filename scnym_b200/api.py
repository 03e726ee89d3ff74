"""`scnym_api(task="train" | "predict")` (reference surface: scnym/api.py:64-716).

Same arguments, validation errors, side effects on `adata` (`uns["scNym_train_results"]`,
`obs[key_added]`, `obs[key_added + "_confidence"]`, `uns["scNym_probabilities"]`,
`obsm["X_scnym"]`) and files (`scnym_train_results.pkl`, checkpoints).  `adata` may be an
`anndata.AnnData` or any object with the same attributes (`scnym_b200.utils.SimpleAnnData`).
Prediction makes ONE pass over the atlas (argmax, probabilities and embedding together) instead
of the reference's two (scnym/api.py:679-708).  Everything outside train/predict
(`scnym_tune`, `scnym_interpret`, `atlas2target`, pretrained-weight download; api.py:719-1523)
is outside the hot path.
"""
import copy
import logging
import os
import os.path as osp
import pickle
from typing import Optional, Union

import numpy as np
import pandas as pd
import torch

from . import main, model, predict, utils  # noqa: F401

logger = logging.getLogger(__name__)

TASKS = ("train", "predict")

CONFIGS = {
    "default": {
        "n_epochs": 100, "patience": 40, "lr": 1.0, "optimizer_name": "adadelta", "weight_decay": 1e-4, "batch_size": 256,
        "balanced_classes": False, "weighted_classes": False, "mixup_alpha": 0.3, "unsup_max_weight": 1.0,
        "unsup_mean_teacher": False, "ssl_method": "mixmatch",
        "ssl_kwargs": {"augment_pseudolabels": False, "augment": "log1p_drop", "unsup_criterion": "mse", "n_augmentations": 1,
                       "T": 0.5, "ramp_epochs": 100, "burn_in_epochs": 0, "dan_criterion": True, "dan_ramp_epochs": 20,
                       "dan_max_weight": 0.1, "min_epochs": 20},
        "model_kwargs": {"n_hidden": 256, "n_layers": 2, "init_dropout": 0.0, "residual": False},
        "tensorboard": False,
    },
}
CONFIGS["no_new_identity"] = copy.deepcopy(CONFIGS["default"])
CONFIGS["no_new_identity"]["description"] = ("Train scNym models with MixMatch and a domain adversary, assuming no new cell types "
                                             "in the target data.")
CONFIGS["new_identity_discovery"] = copy.deepcopy(CONFIGS["default"])
CONFIGS["new_identity_discovery"]["ssl_kwargs"]["pseudolabel_min_confidence"] = 0.9
CONFIGS["new_identity_discovery"]["ssl_kwargs"]["dan_use_conf_pseudolabels"] = True
CONFIGS["new_identity_discovery"]["description"] = ("Train scNym models with MixMatch and a domain adversary, using pseudolabel "
                                                    "thresholding to allow for new cell type discoveries.")
CONFIGS["no_dan"] = copy.deepcopy(CONFIGS["default"])
CONFIGS["no_dan"]["ssl_kwargs"]["dan_max_weight"] = 0.0
CONFIGS["no_dan"]["ssl_kwargs"]["dan_ramp_epochs"] = 1
CONFIGS["no_dan"]["description"] = "Train scNym models with MixMatch but no domain adversary."
CONFIGS["no_ssl"] = copy.deepcopy(CONFIGS["default"])
CONFIGS["no_ssl"]["ssl_kwargs"]["dan_max_weight"] = 0.0
CONFIGS["no_ssl"]["ssl_kwargs"]["dan_ramp_epochs"] = 1
CONFIGS["no_ssl"]["ssl_kwargs"]["unsup_max_weight"] = 0.0
CONFIGS["no_ssl"]["description"] = "Train scNym models without semi-supervision or a domain adversary."

UNLABELED_TOKEN = "Unlabeled"


def scnym_api(adata, task: str = "train", groupby: str = None, domain_groupby: str = None, out_path: str = "./scnym_outputs",
              trained_model: str = None, config: Union[dict, str] = "new_identity_discovery", key_added: str = "scNym",
              copy: bool = False, **kwargs) -> Optional[object]:
    """Train an scNym model on `adata` or predict cell identities with a trained one."""
    if task not in TASKS:
        raise ValueError(f"{task} is not a valid scNym task.\nmust be one of {TASKS}")
    if type(config) == str:
        if config not in CONFIGS.keys():
            raise ValueError(f"{config} is not a predefined configuration.\nmust be one of {CONFIGS.keys()}.")
        config = CONFIGS[config]
    elif type(config) != dict:
        raise TypeError(f"`config` was a {type(config)}, must be dict or str.")
    else:
        # user dictionaries update the global default in place, as in the reference (api.py:278-283)
        dconf = CONFIGS["default"]
        for k in config.keys():
            dconf[k] = config[k]
        config = dconf
    if not osp.exists(out_path):
        os.makedirs(out_path, exist_ok=True)
    config["out_path"], config["groupby"], config["key_added"] = out_path, groupby, key_added
    config["trained_model"], config["domain_groupby"] = trained_model, domain_groupby

    n_genes = adata.shape[1]
    n_unique_genes = len(np.unique(adata.var_names))
    if n_genes != n_unique_genes:
        raise ValueError("Duplicate Genes Error\nNot all genes passed to scNym were unique.\n"
                         f"{n_genes} genes are present but only {n_unique_genes} unique genes were detected.\n"
                         "Please use unique gene names in your input object.\n"
                         "This can be achieved by running `adata.var_names_make_unique()`")
    # log1p(CPM) sanity checks (api.py:326-345)
    x_max = np.max(adata.X) > np.log1p(1e6)
    x_min = np.min(adata.X) < 0.0
    vals = adata.X if type(adata.X) == np.ndarray else adata.X.data
    int_counts = np.all(np.equal(np.mod(vals, 1), 0))
    if x_max or x_min or int_counts:
        raise ValueError("Normalization error\n`adata.X` does not appear to be log(CountsPerMillion+1) normalized data.\n"
                         "Please replace `adata.X` with log1p(CPM) values.\n>>> # starting from raw counts in `adata.X`\n"
                         ">>> sc.pp.normalize_total(adata, target_sum=1e6))\n>>> sc.pp.log1p(adata)")
    def need_cuda():
        if torch.cuda.is_available():
            print("CUDA compute device found.")
        else:
            raise RuntimeError("No CUDA device found: scnym_b200 runs on the GPU only (no CPU fallback).")

    if task == "train":
        if groupby not in adata.obs.columns:
            raise ValueError(f"{groupby} is not a variable in `adata.obs`")
        need_cuda()
        scnym_train(adata=adata, config=config)
    elif task == "predict":
        if trained_model is None:
            raise ValueError("must provide a path to a trained model for prediction.")
        if not os.path.exists(trained_model) and "pretrained_" not in trained_model:
            raise FileNotFoundError("path to the trained model does not exist.")
        config["model_weights"] = trained_model
        need_cuda()
        scnym_predict(adata=adata, config=config)
    return


def scnym_train(adata, config: dict) -> None:
    """Split labeled / "Unlabeled" cells, build the 90/10 split and domains, call `fit_model`,
    store results (scnym/api.py:392-600)."""
    groupby = config["groupby"]
    n_unlabeled = np.sum(adata.obs[groupby] == UNLABELED_TOKEN)
    if n_unlabeled == 0:
        print("No unlabeled data was found.")
        print(f'Did you forget to set some examples as `"{UNLABELED_TOKEN}"`?')
        print("Proceeding with purely supervised training.\n")
        unlabeled_counts = None
        X = utils.get_adata_asarray(adata)
        y = pd.Categorical(np.array(adata.obs[groupby]), categories=np.unique(adata.obs[groupby])).codes
        class_names = np.unique(adata.obs[groupby])
        train_adata = adata
        target_bidx = np.zeros(adata.shape[0], dtype=bool)
    else:
        print(f"{n_unlabeled} unlabeled observations found.")
        print("Using unlabeled data as a target set for semi-supervised, adversarial training.\n")
        target_bidx = np.asarray(adata.obs[groupby] == UNLABELED_TOKEN)
        train_adata = adata[~target_bidx, :]
        target_adata = adata[target_bidx, :]
        print("training examples: ", train_adata.shape)
        print("target   examples: ", target_adata.shape)
        X = utils.get_adata_asarray(train_adata)
        y = pd.Categorical(np.array(train_adata.obs[groupby]), categories=np.unique(train_adata.obs[groupby])).codes
        unlabeled_counts = utils.get_adata_asarray(target_adata)
        class_names = np.unique(train_adata.obs[groupby])
    y = np.asarray(y, dtype=np.int64)
    print("X: ", X.shape)
    print("y: ", y.shape)
    if "scNym_split" not in adata.obs_keys():
        traintest_idx = np.random.choice(X.shape[0], size=int(np.floor(0.9 * X.shape[0])), replace=False)
        val_idx = np.setdiff1d(np.arange(X.shape[0]), traintest_idx)
    else:
        train_idx = np.where(train_adata.obs["scNym_split"] == "train")[0]
        test_idx = np.where(train_adata.obs["scNym_split"] == "test")[0]
        val_idx = np.where(train_adata.obs["scNym_split"] == "val")[0]
        if len(train_idx) < 100 or len(test_idx) < 10 or len(val_idx) < 10:
            raise RuntimeError("Few samples in user provided data split.\n"
                               f"{len(train_idx)} training samples.\n{len(test_idx)} testing samples.\n"
                               f"{len(val_idx)} validation samples.\nHalting.")
        traintest_idx = (train_idx, test_idx)
    if config.get("domain_groupby", None) is not None:
        domain_groupby = config["domain_groupby"]
        if domain_groupby not in adata.obs.columns:
            raise ValueError(f"no column `{domain_groupby}` exists in `adata.obs`.\n"
                             "if domain labels are specified, a matching column must exist.")
        domains = np.array(pd.Categorical(adata.obs[domain_groupby], categories=np.unique(adata.obs[domain_groupby])).codes,
                           dtype=np.int32)
        input_domain, unlabeled_domain = domains[~target_bidx], domains[target_bidx]
        print("Using user provided domain labels.")
        print(f"Found {len(np.unique(input_domain))} source domains and {len(np.unique(unlabeled_domain))} target domains.")
    else:
        input_domain = unlabeled_domain = None
    if config["trained_model"] is None:
        pretrained = None
    elif "pretrained_" in config["trained_model"]:
        raise NotImplementedError("pretrained model fetching is not supported for training.")
    else:
        pretrained = osp.join(config["trained_model"], "00_best_model_weights.pkl")
        if not osp.exists(pretrained):
            raise FileNotFoundError(f"{pretrained} file not found.")
    acc, loss = main.fit_model(
        X=X, y=y, traintest_idx=traintest_idx, val_idx=val_idx, batch_size=config["batch_size"], n_epochs=config["n_epochs"],
        lr=config["lr"], optimizer_name=config["optimizer_name"], weight_decay=config["weight_decay"], ModelClass=model.CellTypeCLF,
        balanced_classes=config["balanced_classes"], weighted_classes=config["weighted_classes"], out_path=config["out_path"],
        mixup_alpha=config["mixup_alpha"], unlabeled_counts=unlabeled_counts, input_domain=input_domain,
        unlabeled_domain=unlabeled_domain, unsup_max_weight=config["unsup_max_weight"], unsup_mean_teacher=config["unsup_mean_teacher"],
        ssl_method=config["ssl_method"], ssl_kwargs=config["ssl_kwargs"], pretrained=pretrained, patience=config.get("patience", None),
        save_freq=config.get("save_freq", None), tensorboard=config.get("tensorboard", False), **config["model_kwargs"])
    results = {
        "model_path": osp.realpath(osp.join(config["out_path"], "00_best_model_weights.pkl")), "final_acc": acc, "final_loss": loss,
        "n_genes": adata.shape[1], "n_cell_types": len(np.unique(y)), "class_names": class_names,
        "gene_names": list(adata.var_names), "model_kwargs": config["model_kwargs"], "traintest_idx": traintest_idx, "val_idx": val_idx,
    }
    if not osp.exists(results["model_path"]):
        raise FileNotFoundError(f"Model path not found: {results['model_path']}")
    adata.uns["scNym_train_results"] = results
    with open(osp.join(config["out_path"], "scnym_train_results.pkl"), "wb") as f:
        pickle.dump(results, f)


@torch.no_grad()
def scnym_predict(adata, config: dict) -> None:
    """Predict identities, confidences, probabilities and the `X_scnym` embedding
    (scnym/api.py:603-716) in a single pass over the device-resident matrix."""
    if "pretrained_" in config["trained_model"]:
        raise NotImplementedError("Pretrained Request Error\nPretrained weights are no longer supported in scNym.\n")
    with open(osp.join(config["trained_model"], "scnym_train_results.pkl"), "rb") as f:
        results = pickle.load(f)
    P = predict.Predicter(model_weights=osp.join(config["trained_model"], "00_best_model_weights.pkl"), n_genes=results["n_genes"],
                          n_cell_types=results["n_cell_types"], labels=results["class_names"], **config["model_kwargs"])
    print(f"Loaded model predicting {results['n_cell_types']} classes from {results['n_genes']} features")
    print(results["class_names"])
    print("Building a classification matrix...")
    # scnym/api.py:670-675 builds a re-indexed copy of the matrix on the host; here only the column map is built on
    # the host and every row chunk is re-indexed in HBM in front of its first layer (predict.EvalRunner.run)
    X = utils.get_adata_asarray(adata)
    model_genes, sample_genes = np.array(results["gene_names"]), np.array(adata.var_names)
    col_map = None
    if len(model_genes) == len(sample_genes) and np.all(model_genes == sample_genes):
        print("Gene names match exactly, returning input.")
    else:
        col_map = utils.gene_column_map(model_genes, sample_genes)
        print("Found %d common genes." % int((col_map >= 0).sum()))
    print("Predicting cell types...")
    pred, prob, _, emb = P.predict_device(X, want_prob=True, want_embed=True, col_map=col_map)
    pred = pred.cpu().numpy()
    names = [results["class_names"][p] for p in pred]
    prob = pd.DataFrame(prob.cpu().numpy(), columns=results["class_names"], index=adata.obs_names)
    adata.obs[config["key_added"]] = names
    adata.obs[config["key_added"] + "_confidence"] = np.max(prob, axis=1)
    adata.uns["scNym_probabilities"] = prob
    adata.obsm["X_scnym"] = emb.cpu().numpy()
