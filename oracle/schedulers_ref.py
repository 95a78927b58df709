"""CPU oracle: restatement of the diffusers==0.23.0 schedulers the reference drives
(SURVEY.md A.6).  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (third-party, no golden vectors).

Call sites in the reference: ``scheduler.set_timesteps`` (pipline_StableDiffusion_ConsistentID.py:510),
``scale_model_input`` (:540), ``step(...).prev_sample`` (:569-571), ``init_noise_sigma`` (prepare_latents),
``add_noise`` (pipelines/StableDIffusionControlNetInpaint_ConsistentID.py:446).
Scripts select Euler (infer.py:33) and DDIM (demo/controlnet_demo.py:67).
SD config: scaled_linear betas 0.00085..0.012, 1000 train steps, steps_offset=1, "leading" spacing, epsilon prediction.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


def sd_alphas_cumprod(n_train=1000, beta_start=0.00085, beta_end=0.012):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n_train, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class _Base:
    order = 1

    def __init__(self, n_train=1000, steps_offset=1):
        self.alphas_cumprod = sd_alphas_cumprod(n_train)
        self.n_train = n_train
        self.steps_offset = steps_offset
        self.config = SimpleNamespace(num_train_timesteps=n_train, steps_offset=steps_offset,
                                      timestep_spacing="leading", prediction_type="epsilon")
        self.timesteps = None

    def _leading(self, n):
        ratio = self.n_train // n
        return (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset

    def add_noise(self, x0, noise, timesteps):
        a = self.alphas_cumprod.to(x0.device)[timesteps].to(x0.dtype)
        while a.ndim < x0.ndim:
            a = a[..., None]
        return a ** 0.5 * x0 + (1 - a) ** 0.5 * noise


class DDIMRef(_Base):
    init_noise_sigma = 1.0

    def set_timesteps(self, n, device=None):
        self.num_inference_steps = n
        self.timesteps = torch.from_numpy(self._leading(n)).to(device)

    def scale_model_input(self, x, t=None):
        return x

    def step(self, eps, t, x, eta=0.0, **kw):
        t = int(t)
        prev_t = t - self.n_train // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.alphas_cumprod[0]   # set_alpha_to_one=False
        b_t = 1 - a_t
        x0 = (x - b_t ** 0.5 * eps) / a_t ** 0.5
        direction = (1 - a_p) ** 0.5 * eps
        return SimpleNamespace(prev_sample=a_p ** 0.5 * x0 + direction, pred_original_sample=x0)


class EulerRef(_Base):
    def set_timesteps(self, n, device=None):
        self.num_inference_steps = n
        ts = self._leading(n).astype(np.float32)
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32)).to(device)
        self.timesteps = torch.from_numpy(ts).to(device)
        self._step_index = None

    @property
    def init_noise_sigma(self):
        return float((self.sigmas.max() ** 2 + 1) ** 0.5)       # "leading" spacing

    def _index(self, t):
        return int((self.timesteps == float(t)).nonzero()[0].item())

    def scale_model_input(self, x, t):
        s = self.sigmas[self._index(t)]
        return x / ((s ** 2 + 1) ** 0.5)

    def add_noise(self, x0, noise, timesteps):
        """EulerDiscreteScheduler.add_noise (diffusers 0.23): x0 + sigma[index of t in self.timesteps] * noise."""
        sig = torch.stack([self.sigmas[self._index(t)] for t in torch.as_tensor(timesteps).reshape(-1)]).to(x0.dtype)
        while sig.ndim < x0.ndim:
            sig = sig[..., None]
        return x0 + noise * sig

    def step(self, eps, t, x, **kw):
        i = self._index(t)
        s, s_next = self.sigmas[i], self.sigmas[i + 1]
        x0 = x - s * eps
        d = (x - x0) / s
        return SimpleNamespace(prev_sample=x + d * (s_next - s), pred_original_sample=x0)


class DPMSolverPP2MRef(_Base):
    """dpmsolver++ multistep, solver_order=2, midpoint, lower_order_final=True."""
    init_noise_sigma = 1.0

    def set_timesteps(self, n, device=None):
        self.num_inference_steps = n
        # DPMSolverMultistepScheduler.set_timesteps, "leading": step_ratio = T // (n + 1); (arange(n + 1) * ratio).round()[::-1][:-1] + offset
        r1 = self.n_train // (n + 1)
        self.timesteps = torch.from_numpy((np.arange(0, n + 1) * r1).round()[::-1][:-1].copy().astype(np.int64) + self.steps_offset).to(device)
        a = self.alphas_cumprod
        self.alpha_t, self.sigma_t = a ** 0.5, (1 - a) ** 0.5
        self.lambda_t = torch.log(self.alpha_t) - torch.log(self.sigma_t)
        self._x0_prev, self._t_prev, self._lower = None, None, 0

    def scale_model_input(self, x, t=None):
        return x

    def step(self, eps, t, x, **kw):
        t = int(t)
        idx = int((self.timesteps == t).nonzero()[0].item())
        last = idx == len(self.timesteps) - 1
        prev_t = 0 if last else int(self.timesteps[idx + 1])
        x0 = (x - self.sigma_t[t] * eps) / self.alpha_t[t]
        lam_p, lam_t = self.lambda_t[prev_t], self.lambda_t[t]
        h = lam_p - lam_t
        a_p, s_p, s_t = self.alpha_t[prev_t], self.sigma_t[prev_t], self.sigma_t[t]
        first_order = self._lower < 1 or (last and len(self.timesteps) < 15)
        if first_order:
            out = (s_p / s_t) * x - a_p * (torch.exp(-h) - 1.0) * x0
        else:
            h0 = lam_t - self.lambda_t[self._t_prev]
            r0 = h0 / h
            d1 = (1.0 / r0) * (x0 - self._x0_prev)
            out = (s_p / s_t) * x - a_p * (torch.exp(-h) - 1.0) * x0 - 0.5 * a_p * (torch.exp(-h) - 1.0) * d1
        self._x0_prev, self._t_prev = x0, t
        self._lower = min(self._lower + 1, 2)
        return SimpleNamespace(prev_sample=out, pred_original_sample=x0)


def make_scheduler(name):
    return {"ddim": DDIMRef, "euler": EulerRef, "dpmpp2m": DPMSolverPP2MRef}[name]()
