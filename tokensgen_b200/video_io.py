"""Input / output glue of the CLI (host I/O): `load_video` with the sampling, centre-crop and [-1, 1]
scaling of longvgen/data/long_video.py:28-76 (decord is replaced by OpenCV, which this image has), and mp4 export."""
from __future__ import annotations

import numpy as np
import torch


def _center_crop_resize(frames: torch.Tensor, out_hw) -> torch.Tensor:
    """resize_for_rectangle_crop(..., reshape_mode="center") (longvgen/data/utils.py:112-141): scale so the clip covers the
    target — BICUBIC, the resized side TRUNCATED with int() exactly as the reference computes it, antialiased as
    torchvision's tensor `resize` does (0.19: antialias defaults to True; float tensors are not clamped) — then centre crop."""
    th, tw = out_hw
    _, _, h, w = frames.shape
    if w / h > tw / th:
        nh, nw = th, int(w * th / h)
    else:
        nh, nw = int(h * tw / w), tw
    frames = torch.nn.functional.interpolate(frames, size=(nh, nw), mode="bicubic", align_corners=False, antialias=True)
    top, left = (nh - th) // 2, (nw - tw) // 2
    return frames[:, :, top:top + th, left:left + tw]


def load_video(video_path, output_res, nf_per_chunk, pad_to_fit, sample_fps, start_t, end_t, max_num_chunks, crop_to_fit=False):
    import cv2
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise IOError(f"cannot open {video_path}")
    fps, total = cap.get(cv2.CAP_PROP_FPS), int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    sample_fps = fps if sample_fps == -1 else sample_fps
    end_t = total / fps if end_t == -1 else min(total / fps, end_t)
    assert 0 <= start_t < end_t and sample_fps > 0
    idx = np.linspace(int(start_t * fps), int(end_t * fps), int((end_t - start_t) * sample_fps), endpoint=False).astype(int)
    num_chunks = min(len(idx) // nf_per_chunk, max_num_chunks)
    idx = idx[:num_chunks * nf_per_chunk]
    assert len(idx) > 0, "sample_idx is empty!"
    wanted, frames, pos = set(idx.tolist()), {}, 0
    while pos <= idx[-1]:
        ok, img = cap.read()
        if not ok:
            break
        if pos in wanted:
            frames[pos] = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        pos += 1
    cap.release()
    missing = [i for i in idx.tolist() if i not in frames]
    if missing:
        raise IOError(f"{video_path}: the decoder ended at frame {pos} but frame {missing[0]} was requested "
                      f"(container reports {total} frames at {fps:.3f} fps)")
    video = torch.from_numpy(np.stack([frames[i] for i in idx.tolist()])).float().permute(0, 3, 1, 2)   # f c h w
    if crop_to_fit:
        px = _center_crop_resize(video / 255.0, tuple(output_res)) * 2 - 1
    else:
        if tuple(video.shape[-2:]) != tuple(output_res):
            raise NotImplementedError("pad_to_fit resizing is not used by the shipped configs (crop_to_fit: true)")
        px = video / 127.5 - 1.0
    return px.unsqueeze(0)


def export_to_video(frames, path: str, fps: int = 10) -> str:
    """frames: [F, H, W, 3] float in [0, 1] (numpy) or a list of PIL images."""
    import cv2
    arr = np.stack([np.asarray(f) for f in frames]) if isinstance(frames, (list, tuple)) else np.asarray(frames)
    if arr.dtype != np.uint8:
        arr = (arr.clip(0, 1) * 255).round().astype(np.uint8)
    h, w = arr.shape[1:3]
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (w, h))
    for f in arr:
        wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    wr.release()
    return path
