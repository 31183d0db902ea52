"""Input pipeline (fusion_gcn_b200/pipeline.py) on CPU: same on-disk format, same batches in the same order as the
reference's DataLoader(MultiModalDataset) (torch_src/dataset.py:15-58, torch_src/loader.py:22-33, session/training.py:18-25)."""
import os

import numpy as np
import pytest
import torch

from fusion_gcn_b200.pipeline import FeatureStore, PrefetchLoader


def write_split(root, split, n, shape=(2, 6, 5, 3), modalities=("skeleton",), dtype=np.float32, seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(root, exist_ok=True)
    for m in modalities:
        np.save(os.path.join(root, f"{m}_{split}_features.npy"), rng.standard_normal((n,) + shape).astype(dtype))
    np.save(os.path.join(root, f"{split}_labels.npy"), rng.integers(0, 7, n))


def reference_batches(root, split, batch, shuffle, drop_last, seed):
    from oracle import ref_loader
    ref_loader.load()
    from dataset import MultiModalDataset          # the reference's own classes (torch_src/ on sys.path)
    from loader import NumpyDatasetLoader
    from torch.utils.data import DataLoader
    ds = MultiModalDataset([(root, NumpyDatasetLoader())], split)
    torch.manual_seed(seed)
    return [(f, l, i) for f, l, i in DataLoader(ds, batch, shuffle=shuffle, drop_last=drop_last)], ds


@pytest.mark.parametrize("shuffle,drop_last", [(True, True), (False, False)])
@pytest.mark.parametrize("modalities,dtype", [(("skeleton",), np.float32), (("skeleton", "imu"), np.float64)])
def test_same_batches_as_the_reference_loader(tmp_path, shuffle, drop_last, modalities, dtype):
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not available")
    root = str(tmp_path)
    write_split(root, "train", 23, modalities=modalities, dtype=dtype)
    want, ds = reference_batches(root, "train", 4, shuffle, drop_last, seed=11)
    store = FeatureStore(root, "train")
    assert store.get_input_shape() == {k: tuple(v) for k, v in ds.get_input_shape().items()} and store.get_num_classes() == ds.get_num_classes()
    torch.manual_seed(11)
    got = list(PrefetchLoader(store, 4, shuffle=shuffle, drop_last=drop_last, device="cpu"))
    assert len(got) == len(want) == len(PrefetchLoader(store, 4, shuffle=shuffle, drop_last=drop_last, device="cpu"))
    for (f, l, i), (fr, lr, ir) in zip(got, want):
        assert torch.equal(i, ir) and torch.equal(l, lr.long())
        if isinstance(fr, dict):
            assert set(f) == set(fr)
            for k in fr:
                assert f[k].dtype == torch.float32 and torch.equal(f[k], fr[k].float())
        else:
            assert f.dtype == torch.float32 and torch.equal(f, fr.float())


def test_adopts_a_reference_dataset_and_two_epochs_differ(tmp_path):
    root = str(tmp_path)
    write_split(root, "val", 10)
    store = FeatureStore(root, "val", in_memory=True)

    class FakeDataset:                       # the duck type of MultiModalDataset the drop-in launcher hands over
        labels_data = store.labels
        features_data = {"skeleton": (None, store.features["skeleton"])}
    loader = PrefetchLoader(FakeDataset(), 3, shuffle=True, device="cpu")
    torch.manual_seed(0)
    e1 = torch.cat([i for _, _, i in loader])
    e2 = torch.cat([i for _, _, i in loader])
    assert sorted(e1.tolist()) == sorted(e2.tolist()) == list(range(10)) and not torch.equal(e1, e2)
    assert len(loader.dataset) == 10


def test_feature_store_from_arrays_feeds_the_loader_in_order():
    """FeatureStore.from_arrays (what bench.py's end-to-end leg builds its synthetic epoch from): sequential batches come back in
    sample order with the labels that belong to them; a short feature array is rejected."""
    from fusion_gcn_b200.pipeline import FeatureStore, PrefetchLoader
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((10, 2, 6, 5, 3)).astype(np.float32)
    labels = np.arange(10, dtype=np.int64)
    store = FeatureStore.from_arrays({"skeleton": feats}, labels)
    assert len(store) == 10 and store.get_input_shape() == {"skeleton": (2, 6, 5, 3)}
    seen = []
    for x, y, idx in PrefetchLoader(store, 4, shuffle=False, drop_last=False, device="cpu"):
        assert torch.equal(x, torch.from_numpy(feats[idx.numpy()])) and torch.equal(y, torch.from_numpy(labels[idx.numpy()]))
        seen += idx.tolist()
    assert seen == list(range(10))
    with pytest.raises(ValueError):
        FeatureStore.from_arrays({"skeleton": feats[:5]}, labels)
