"""diffusers.video_processor.VideoProcessor.postprocess_video as the reference calls it (pipeline_cogvideox_mp_fifo.py:363):
denormalise to [0, 1]; "pt" -> [B, F, C, H, W], "np" -> [B, F, H, W, C] float32."""
import torch


class VideoProcessor:
    def __init__(self, vae_scale_factor=8):
        self.vae_scale_factor = vae_scale_factor

    def postprocess_video(self, video, output_type="np"):
        v = (video / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return v.permute(0, 2, 1, 3, 4)
        if output_type == "np":
            return v.permute(0, 2, 3, 4, 1).float().cpu().numpy()
        raise ValueError(output_type)
