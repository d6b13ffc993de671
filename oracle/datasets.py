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


class TeacherVideoSet(VideoSet):
    """Items that also carry teacher (CLIP) frame features: (feat, teacher_feat, idx, id), the 4-field layout
    collate_frame_val accepts (method/data_provider.py:144-148)."""

    def __init__(self, feats, teacher, prefix="vid"):
        super().__init__(feats, prefix)
        self.teacher = teacher

    def __getitem__(self, i):
        return self.feats[i], self.teacher[i], i, self.ids[i]


class TeacherQuerySet(QuerySet):
    """(feat, teacher_text (1, Dt), idx, id): collate_text_val concatenates the teacher vectors (data_provider.py:159-160)."""

    def __init__(self, feats, teacher, n_videos, prefix="vid"):
        super().__init__(feats, n_videos, prefix)
        self.teacher = teacher

    def __getitem__(self, i):
        return self.feats[i], self.teacher[i], i, self.ids[i]
