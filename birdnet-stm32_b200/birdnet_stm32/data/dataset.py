"""File discovery for class-structured datasets (reference `data/dataset.py:49-99`, without tf.io.gfile)."""

from __future__ import annotations

import os

import numpy as np

SUPPORTED_AUDIO_EXTS = (".wav", ".mp3", ".flac", ".ogg", ".m4a")
_NOISE = {"noise", "silence", "background", "other"}


def load_file_paths_from_directory(directory: str, classes: list[str] | None = None, max_samples: int | None = None,
                                   exts: tuple = SUPPORTED_AUDIO_EXTS) -> tuple[list[str], list[str]]:
    """(globally shuffled file paths, sorted class names without noise-like folders).

    Class = parent directory name; `classes` restricts the walk; `max_samples` caps files per class
    with a uniform random subset; shuffling uses numpy's global RNG like the reference (`:94`).
    """
    by_class: dict[str, list[str]] = {}
    for root, _dirs, names in os.walk(directory):
        label = os.path.basename(root)
        if classes is not None and label not in classes:
            continue
        hits = [os.path.join(root, n) for n in names if n.lower().endswith(exts)]
        if hits:
            by_class.setdefault(label, []).extend(hits)
    paths: list[str] = []
    for members in by_class.values():
        if max_samples is not None and max_samples > 0 and len(members) > max_samples:
            keep = np.random.permutation(len(members))[:max_samples]
            members = [members[i] for i in keep]
        paths.extend(members)
    np.random.shuffle(paths)
    return paths, sorted(c for c in by_class if c.lower() not in _NOISE)
