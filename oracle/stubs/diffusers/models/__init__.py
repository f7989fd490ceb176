def __getattr__(name):  # `from diffusers.models import AutoencoderKLCogVideoX`: the reference tree carries its own copy
    if name == "AutoencoderKLCogVideoX":
        from longvgen.models.autoencoder_kl_cogvideox import AutoencoderKLCogVideoX
        return AutoencoderKLCogVideoX
    raise AttributeError(name)
