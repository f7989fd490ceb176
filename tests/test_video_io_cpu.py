"""load_video's crop_to_fit path against the reference's own `resize_for_rectangle_crop` (longvgen/data/utils.py:112-141;
ADVICE r1 medium): BICUBIC, int() truncation of the resized side, centre crop.  The reference function is pure
torchvision; it is imported from the reference tree where that is present (build container) and restated otherwise."""
import importlib.util
import os

import numpy as np
import pytest
import torch


def _reference_fn():
    path = "/root/reference/longvgen/data/utils.py"
    if os.path.exists(path):
        spec = importlib.util.spec_from_file_location("ref_data_utils", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.resize_for_rectangle_crop, "reference"

    def restated(arr, image_size, reshape_mode="center"):      # utils.py:112-141, "center" branch
        from torchvision.transforms import InterpolationMode
        from torchvision.transforms.functional import crop, resize
        if arr.shape[3] / arr.shape[2] > image_size[1] / image_size[0]:
            arr = resize(arr, size=[image_size[0], int(arr.shape[3] * image_size[0] / arr.shape[2])], interpolation=InterpolationMode.BICUBIC)
        else:
            arr = resize(arr, size=[int(arr.shape[2] * image_size[1] / arr.shape[3]), image_size[1]], interpolation=InterpolationMode.BICUBIC)
        h, w = arr.shape[2], arr.shape[3]
        return crop(arr, top=(h - image_size[0]) // 2, left=(w - image_size[1]) // 2, height=image_size[0], width=image_size[1])
    return restated, "restated"


@pytest.mark.parametrize("h,w", [(480, 720), (720, 1280), (1080, 1920), (487, 853), (853, 487), (360, 641), (1000, 1000)])
def test_center_crop_resize_equals_the_reference_function(h, w):
    pytest.importorskip("torchvision")
    from tokensgen_b200.video_io import _center_crop_resize
    fn, kind = _reference_fn()
    g = torch.Generator().manual_seed(h * 7 + w)
    frames = torch.rand(3, 3, h, w, generator=g)
    want = fn(frames.clone(), (480, 720), reshape_mode="center")
    got = _center_crop_resize(frames, (480, 720))
    assert tuple(got.shape) == tuple(want.shape) == (3, 3, 480, 720), kind
    assert torch.equal(got, want), (kind, (got - want).abs().max().item())


def test_load_video_reports_a_short_stream(tmp_path, monkeypatch):
    """cap.read() ending before the last requested index is an IOError naming the frame, not a KeyError."""
    cv2 = pytest.importorskip("cv2")
    from tokensgen_b200.video_io import export_to_video, load_video
    path = str(tmp_path / "short.mp4")
    export_to_video(np.random.RandomState(0).rand(12, 64, 96, 3).astype(np.float32), path, fps=8)
    cap = cv2.VideoCapture(path)
    n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    cap.release()
    v = load_video(path, (64, 96), 4, False, -1, 0, -1, 8, crop_to_fit=False)
    assert tuple(v.shape) == (1, (n // 4) * 4, 3, 64, 96) and v.min() >= -1 and v.max() <= 1
    real = cv2.VideoCapture

    class Overcounting:                      # a container whose header promises twice the frames the stream holds
        def __init__(self, p):
            self.c = real(p)

        def get(self, prop):
            v = self.c.get(prop)
            return v * 2 if prop == cv2.CAP_PROP_FRAME_COUNT else v

        def __getattr__(self, name):
            return getattr(self.c, name)

    monkeypatch.setattr(cv2, "VideoCapture", Overcounting)
    with pytest.raises(IOError, match="was requested"):
        load_video(path, (64, 96), 4, False, -1, 0, -1, 8, crop_to_fit=False)
