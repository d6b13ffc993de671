"""In-memory datasets with the item layout of the reference's VisDataSet4DLDKD / TxtDataSet4DLDKD
(method/data_provider.py:309,354): (feat, idx, id).  Shared by the oracle, the tests and bench.py."""
import torch.utils.data as data


class VideoSet(data.Dataset):
    def __init__(self, feats, prefix="vid"):
        self.feats = feats
        self.ids = [f"{prefix}{n}" for n in range(len(feats))]

    def __len__(self):
        return len(self.feats)

    def __getitem__(self, i):
        return self.feats[i], i, self.ids[i]


class QuerySet(data.Dataset):
    """Caption ids 'vid{q mod Nv}#enc#{q div Nv}' so that get_gt (method/eval.py:43-57) links them."""

    def __init__(self, feats, n_videos, prefix="vid"):
        self.feats = feats
        self.ids = [f"{prefix}{q % n_videos}#enc#{q // n_videos}" for q in range(len(feats))]

    def __len__(self):
        return len(self.feats)

    def __getitem__(self, i):
        return self.feats[i], i, self.ids[i]
