"""Host logic: B200Scheduler coefficient tables reproduce the oracle's restatement of the diffusers schedulers."""
import pytest
import torch

sched = pytest.importorskip("consistentid_b200.scheduler")
from oracle.schedulers_ref import make_scheduler


@pytest.mark.parametrize("kind", ["ddim", "euler", "dpmpp2m"])
@pytest.mark.parametrize("steps", [4, 20, 30])
def test_scheduler_matches_oracle(kind, steps):
    ref = make_scheduler(kind)
    ref.set_timesteps(steps)
    ours = sched.B200Scheduler(kind)
    ours.set_timesteps(steps)
    assert [float(t) for t in ref.timesteps] == [float(t) for t in ours.timesteps]
    assert abs(float(ref.init_noise_sigma) - float(ours.init_noise_sigma)) < 1e-5
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64) * float(ref.init_noise_sigma)
    xr, xo = x.clone(), x.clone()
    for t in ref.timesteps:
        eps = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
        a = ref.scale_model_input(xr, t)
        b = ours.scale_model_input(xo, t)
        assert torch.allclose(a, b.double(), rtol=2e-4, atol=2e-4)
        xr = ref.step(eps, t, xr).prev_sample
        xo = ours.step(eps.float(), t, xo.float()).prev_sample.double()
        assert torch.allclose(xr, xo, rtol=2e-4, atol=2e-4), (kind, float(t), (xr - xo).abs().max())


def test_add_noise():
    ref = make_scheduler("ddim")
    ours = sched.B200Scheduler("ddim")
    x0, n = torch.randn(2, 4, 4, 4), torch.randn(2, 4, 4, 4)
    t = torch.tensor([10, 500])
    assert torch.allclose(ref.add_noise(x0, n, t), ours.add_noise(x0, n, t), atol=1e-5)


def test_add_noise_euler():
    ref = make_scheduler("euler"); ours = sched.B200Scheduler("euler")
    ref.set_timesteps(20); ours.set_timesteps(20)
    x0, n = torch.randn(2, 4, 4, 4), torch.randn(2, 4, 4, 4)
    t = ref.timesteps[[3, 17]]
    assert torch.allclose(ref.add_noise(x0, n, t), ours.add_noise(x0, n, t), rtol=1e-5, atol=1e-5)


def test_published_schedule_constants():
    """Known-answer anchors of the Stable Diffusion noise schedule (scaled_linear betas 0.00085..0.012, 1000 steps) that every SD
    implementation publishes: alpha_bar_0 = 0.99915, alpha_bar_999 = 0.00466, sigma_max = 14.6146, sigma_min = 0.0292."""
    from oracle.schedulers_ref import sd_alphas_cumprod
    for acp in (sched.alphas_cumprod(), sd_alphas_cumprod().double().numpy()):
        assert abs(acp[0] - 0.99915) < 1e-5 and abs(acp[999] - 0.00466) < 1e-5
        assert abs(((1 - acp[999]) / acp[999]) ** 0.5 - 14.6146) < 1e-3
        assert abs(((1 - acp[0]) / acp[0]) ** 0.5 - 0.0292) < 1e-4
    s = sched.B200Scheduler("euler")
    s.set_timesteps(1000)                       # leading spacing + steps_offset 1: t = 1000 .. 1, the first one clamps to the table end
    assert abs(s.sigmas[0] - 14.6146) < 1e-3 and s.sigmas[-1] == 0.0 and abs(s.init_noise_sigma - (14.6146 ** 2 + 1) ** 0.5) < 1e-3
    d = sched.B200Scheduler("ddim")
    d.set_timesteps(50)
    assert list(d.timesteps[:3].tolist()) == [981, 961, 941] and int(d.timesteps[-1]) == 1     # leading spacing, steps_offset 1
