"""TEST INFRASTRUCTURE — seeded synthetic feature banks shared by the golden generator,
the tests and bench.py."""
import torch


def synth_banks(seed, C, Dv, D, n_img, txt_per_class, n_val):
    """Learnable synthetic banks: class prototypes + noise, separate prototype sets per
    modality so the text rows help but do not coincide with the image rows."""
    g = torch.Generator().manual_seed(seed)
    mu_i = torch.randn(C, Dv, generator=g)
    mu_t = torch.randn(C, D, generator=g)
    yi = torch.randint(0, C, (n_img,), generator=g)
    xi = mu_i[yi] + 1.0 * torch.randn(n_img, Dv, generator=g)
    yt = torch.arange(C).repeat_interleave(txt_per_class)
    xt = mu_t[yt] + 0.5 * torch.randn(C * txt_per_class, D, generator=g)
    yv = torch.randint(0, C, (n_val,), generator=g)
    xv = mu_i[yv] + 1.0 * torch.randn(n_val, Dv, generator=g)
    return xi, yi, xt, yt, xv, yv
