"""Golden vectors for the GAN-phase discriminators and losses (VocosDiscriminator.forward_disc / forward_gen), generated from
the REAL reference modules.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_disc.py

Weights come from oracle.spec.deterministic_state_dict over oracle.discriminators.discriminator_shapes() (41.7 M parameters
are not stored); tests/golden/discriminator.npz holds the inputs, the loss terms, the gradient of the generator-side loss with
respect to wav_hat (norm + a strided slice), output sizes and per-feature-map means."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    MG.install_stubs()
    sys.path.insert(0, MG.REF)
    sys.path.append(ROOT)
    from oracle.discriminators import discriminator_shapes
    from oracle.spec import ModelSpec, deterministic_state_dict

    import optispeech  # noqa: F401
    assert optispeech.__file__.startswith(MG.REF)
    spec = ModelSpec()
    _, fe = MG.build_reference_generator(spec)
    disc = MG.build_reference_discriminator(spec, fe)
    shapes = discriminator_shapes()
    ref_shapes = {k: tuple(v.shape) for k, v in disc.state_dict().items() if k.startswith(("multiperioddisc", "multiresddisc"))}
    assert ref_shapes == shapes, "oracle/discriminators.py shape table is out of date"
    assert sum(p.numel() for p in disc.parameters()) == 41_705_968
    sd = deterministic_state_dict(shapes, seed=0)
    missing, unexpected = disc.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("melspec_loss", "mr_stft_loss")) for k in missing), (missing, unexpected)
    disc.eval()
    g = torch.Generator().manual_seed(77)
    B, Lw = 2, 16384
    t = torch.arange(Lw) / 22050.0
    wav = 0.5 * torch.sin(2 * np.pi * 220.0 * t)[None] * torch.tensor([[1.0], [0.6]]) + 0.05 * torch.randn(B, Lw, generator=g)
    wav = wav.clamp(-1, 1)
    wav_hat = (wav + 0.1 * torch.randn(B, Lw, generator=g)).clamp(-1, 1).requires_grad_(True)
    loss_d, log_d = disc.forward_disc(wav, wav_hat.detach())
    loss_g, log_g = disc.forward_gen(wav, wav_hat)
    loss_g.backward()
    fx = dict(wav=wav.numpy(), wav_hat=wav_hat.detach().numpy(), loss_disc=float(loss_d), loss_gen=float(loss_g),
              dwav_hat_norm=float(wav_hat.grad.norm()), dwav_hat_slice=wav_hat.grad[:, ::64].numpy())
    fx.update({f"disc_{k}": float(v) for k, v in log_d.items()})
    fx.update({f"gen_{k}": float(v) for k, v in log_g.items()})
    with torch.no_grad():
        outs_r, _, fr, _ = disc.multiperioddisc(y=wav, y_hat=wav_hat.detach())
        fx["mpd_out_sizes"] = np.array([o.shape[1] for o in outs_r])
        fx["mpd_fmap_means"] = np.array([[float(f.mean()) for f in fm] for fm in fr])
        outs_r, _, fr, _ = disc.multiresddisc(y=wav, y_hat=wav_hat.detach())
        fx["mrd_out_sizes"] = np.array([o.shape[1] for o in outs_r])
        fx["mrd_fmap_means"] = np.array([[float(f.mean()) for f in fm] for fm in fr])
    np.savez_compressed(os.path.join(HERE, "discriminator.npz"), **fx)
    print("disc", float(loss_d), "gen", float(loss_g), {k: round(float(v), 5) for k, v in log_g.items()})


if __name__ == "__main__":
    main()
