"""Checkpoint / sample I/O with the reference's call signatures (src/utils.py:11-141).

Out of scope for acceleration (pure I/O) but part of the drop-in surface the
entry points call.  One behavioural difference, additive: google-cloud-storage is imported
lazily (the reference imports it at module import, which fails where it is not installed).
``load_checkpoint`` calls ``torch.load(path, map_location=device)`` exactly as the reference does
(torch's default ``weights_only`` applies; both checkpoint formats are plain tensor dicts).
"""
from __future__ import annotations

import contextlib
import os
import tempfile
from pathlib import Path
from typing import Union

import torch

PathLike = Union[str, Path]
_GS = "gs://"


class _LazyStorage:
    """Stands where the reference has ``from google.cloud import storage``: resolves on first use,
    and is patchable as ``utils.storage.Client`` exactly like the reference module attribute."""

    def __getattr__(self, name):
        from google.cloud import storage as real  # noqa: PLC0415  (deliberately lazy)

        return getattr(real, name)


storage = _LazyStorage()


def is_gcs_path(path: PathLike) -> bool:
    return str(path).startswith(_GS)


def parse_gcs_path(gcs_path: str) -> tuple[str, str]:
    if not gcs_path.startswith(_GS):
        raise ValueError(f"Not a GCS path: {gcs_path}")
    bucket, _, blob = gcs_path[len(_GS):].partition("/")
    return bucket, blob


def _blob(gcs_path: str):
    bucket, name = parse_gcs_path(gcs_path)
    return storage.Client().bucket(bucket).blob(name)


def download_from_gcs(gcs_path: str, local_path: str) -> None:
    _blob(gcs_path).download_to_filename(local_path)


def upload_to_gcs(local_path: str, gcs_path: str) -> None:
    _blob(gcs_path).upload_from_filename(local_path)


@contextlib.contextmanager
def _scratch_file(suffix: str, mode: str = "w+b"):
    """A named temp file that is always unlinked, whatever happens inside the block."""
    with tempfile.NamedTemporaryFile(mode=mode, suffix=suffix, delete=False) as tmp:
        try:
            yield tmp
        finally:
            os.unlink(tmp.name)   # still open here: unlinking an open file is fine on POSIX


def load_checkpoint(ckpt_path: PathLike, device: str) -> dict:
    ckpt_path = str(ckpt_path)
    if not is_gcs_path(ckpt_path):
        return torch.load(ckpt_path, map_location=device)
    with _scratch_file(".pth") as tmp:
        try:
            print(f"Downloading checkpoint from GCS: {ckpt_path}")
            download_from_gcs(ckpt_path, tmp.name)
            return torch.load(tmp.name, map_location=device)
        except Exception as e:
            raise RuntimeError(f"Failed to download checkpoint from {ckpt_path}: {e}")


def save_checkpoint(model_state: dict, ckpt_path: PathLike) -> None:
    ckpt_path = str(ckpt_path)
    if not is_gcs_path(ckpt_path):
        torch.save(model_state, ckpt_path)
        print(f"✔ Saved checkpoint to {ckpt_path}")
        return
    with _scratch_file(".pth") as tmp:
        try:
            torch.save(model_state, tmp.name)
            print(f"Uploading checkpoint to GCS: {ckpt_path}")
            upload_to_gcs(tmp.name, ckpt_path)
            print(f"✔ Uploaded checkpoint to {ckpt_path}")
        except Exception as e:
            raise RuntimeError(f"Failed to upload checkpoint to {ckpt_path}: {e}")


def save_samples(content: Union[str, bytes], sample_path: PathLike, mode: str = "w") -> None:
    sample_path = str(sample_path)
    if not is_gcs_path(sample_path):
        target = Path(sample_path)
        target.parent.mkdir(parents=True, exist_ok=True)
        if isinstance(content, str):
            target.write_text(content)
        else:
            target.write_bytes(content)
        print(f"✔ Saved sample to {sample_path}")
        return
    with _scratch_file(Path(sample_path).suffix, mode=mode) as tmp:
        try:
            tmp.write(content)
            tmp.flush()
            tmp.close()
            print(f"Uploading sample to GCS: {sample_path}")
            upload_to_gcs(tmp.name, sample_path)
            print(f"✔ Uploaded sample to {sample_path}")
        except Exception as e:
            raise RuntimeError(f"Failed to upload sample to {sample_path}: {e}")


def get_vertex_checkpoint_path(base_name: str) -> str:
    root = os.environ.get("AIP_MODEL_DIR")
    return os.path.join(root, base_name) if root is not None else base_name


def get_samples_dir(base_dir: str = "samples") -> PathLike:
    root = os.environ.get("AIP_MODEL_DIR")
    if root is None:
        return Path(base_dir)
    if root.startswith(_GS):
        # plain string for gs:// so pathlib cannot collapse the double slash
        return f"{root.rstrip('/')}/{base_dir.strip('/')}"
    return Path(root) / base_dir
