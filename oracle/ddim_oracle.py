"""ORACLE (test infrastructure, not product code): CPU restatement of the DDIM / CFG sampling loop.

Restates, in plain numpy + PyTorch, the arithmetic of
  DDPM.register_schedule            DynamiCrafter/lvdm/models/ddpm3d.py:119-182  (bf16 buffers, :139)
  LatentDiffusion scale_arr          ddpm3d.py:505-510
  predict_{start,eps}_from_z_and_v   ddpm3d.py:235-247
  make_ddim_timesteps / _sampling_parameters / rescale_zero_terminal_snr / rescale_noise_cfg
                                     lvdm/models/utils_diffusion.py:56-91,112-158
  DDIMSampler.make_schedule / ddim_sampling / p_sample_ddim
                                     lvdm/models/samplers/ddim.py:24-63,141-215,218-290
  DiffusionWrapper.forward 'hybrid'  ddpm3d.py:1077-1081
Pinned against the reference's own classes by `oracle/make_golden.py` -> `tests/golden/ddim_*.npz`.
Only tests, smoke() and bench.py's CPU-baseline legs may import this module.
"""
import numpy as np
import torch


def model_schedule(timesteps=1000, linear_start=0.00085, linear_end=0.012,
                   rescale_betas_zero_snr=True, use_dynamic_rescale=True, base_scale=0.7,
                   turning_step=400, **_):
    """The diffusion wrapper's buffers, stored in bf16 exactly like the reference does."""
    # the reference builds betas with torch.linspace in float64 (utils_diffusion.py:33-35)
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
    if rescale_betas_zero_snr:
        abar_sqrt = np.sqrt(np.cumprod(1.0 - betas, axis=0))
        a0, aT = abar_sqrt[0].copy(), abar_sqrt[-1].copy()
        abar_sqrt -= aT
        abar_sqrt *= a0 / (a0 - aT)
        abar = abar_sqrt ** 2
        alphas = np.concatenate([abar[0:1], abar[1:] / abar[:-1]])
        betas = 1 - alphas
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    bf = lambda a: torch.tensor(a, dtype=torch.bfloat16)
    buf = dict(
        num_timesteps=int(timesteps),
        betas=bf(betas),
        alphas_cumprod=bf(ac),
        alphas_cumprod_prev=bf(np.append(1.0, ac[:-1])),
        sqrt_alphas_cumprod=bf(np.sqrt(ac)),
        sqrt_one_minus_alphas_cumprod=bf(np.sqrt(1.0 - ac)),
        use_dynamic_rescale=bool(use_dynamic_rescale),
    )
    if use_dynamic_rescale:
        buf["scale_arr"] = bf(np.concatenate((np.linspace(1.0, base_scale, turning_step),
                                              np.full(timesteps, base_scale))))
    return buf


def ddim_timesteps(method, S, T=1000):
    if method == "uniform":
        return np.asarray(list(range(0, T, T // S))) + 1
    if method == "uniform_trailing":
        return np.flip(np.round(np.arange(T, 0, -T / S))).astype(np.int64) - 1
    if method == "quad":
        return ((np.linspace(0, np.sqrt(T * .8), S)) ** 2).astype(int) + 1
    raise NotImplementedError(method)


def ddim_tables(buf, S, spacing="uniform_trailing", eta=1.0):
    """Per-index scalars exactly as p_sample_ddim will read them (as Python floats)."""
    ts = ddim_timesteps(spacing, S, buf["num_timesteps"])
    ac = buf["alphas_cumprod"].to(torch.float32)
    alphas = ac[ts]                                                    # fp32 tensor
    alphas_prev = np.asarray([float(ac[0])] + ac[ts[:-1]].tolist())    # float64 ndarray
    # `ndarray / Tensor` dispatches to Tensor.__rtruediv__ = reciprocal(self) * other, so the
    # reference evaluates 1/(1-a_t) in fp32 before promoting to fp64 (utils_diffusion.py:86).
    a32 = alphas.numpy()
    recip = (np.float32(1.0) / (np.float32(1.0) - a32)).astype(np.float64)
    sigmas = eta * np.sqrt(recip * (1 - alphas_prev) * (1 - a32 / alphas_prev))
    tab = dict(
        timesteps=ts,
        alphas=[float(v) for v in alphas],
        alphas_prev=[float(v) for v in alphas_prev],
        sigmas=[float(v) for v in np.asarray(sigmas, dtype=np.float64)],
        sqrt_one_minus_alphas=[float(v) for v in torch.sqrt(1.0 - alphas)],
    )
    if buf["use_dynamic_rescale"]:
        sa = buf["scale_arr"][ts]
        sp = torch.cat([sa[0:1], sa[:-1]])
        tab["scale"] = [float(v) for v in sa]
        tab["scale_prev"] = [float(v) for v in sp]
    return tab


def rescale_noise_cfg(cfg_out, cond_out, phi):
    dims = list(range(1, cond_out.ndim))
    std_c = cond_out.std(dim=dims, keepdim=True)   # unbiased, torch default
    std = cfg_out.std(dim=dims, keepdim=True)
    return phi * (cfg_out * (std_c / std)) + (1 - phi) * cfg_out


def ddim_step(buf, tab, x, v_cond, v_uncond, step, index, noise, cfg_scale=1.0,
              guidance_rescale=0.0, temperature=1.0, v_uncond_img=None, cfg_img=None):
    """ddim.py:218-290 for parameterization 'v'. All tensor arithmetic happens in x.dtype.
    With `v_uncond_img` it is ddim_multiplecond.py:211-283 (three-branch combine :234, cfg_img defaults to the
    text scale :222-223); the rest of that function is the same arithmetic."""
    b, dt = x.shape[0], x.dtype
    shp = (b,) + (1,) * (x.dim() - 1)
    full = lambda val: torch.full(shp, val, dtype=dt, device=x.device)
    if v_uncond is None or cfg_scale == 1.0:
        v = v_cond
    else:
        if v_uncond_img is None:
            v = v_uncond + cfg_scale * (v_cond - v_uncond)
        else:
            s_img = cfg_scale if cfg_img is None else cfg_img
            v = v_uncond + s_img * (v_uncond_img - v_uncond) + cfg_scale * (v_cond - v_uncond_img)
        if guidance_rescale > 0.0:
            v = rescale_noise_cfg(v, v_cond, guidance_rescale)
    t = torch.full((b,), int(step), dtype=torch.long, device=x.device)
    sa = buf["sqrt_alphas_cumprod"].to(x.device).gather(-1, t).reshape(shp)
    so = buf["sqrt_one_minus_alphas_cumprod"].to(x.device).gather(-1, t).reshape(shp)
    e_t = sa * v + so * x
    pred_x0 = sa * x - so * v
    a_prev, sigma = full(tab["alphas_prev"][index]), full(tab["sigmas"][index])
    if buf["use_dynamic_rescale"]:
        pred_x0 = pred_x0 * (full(tab["scale_prev"][index]) / full(tab["scale"][index]))
    dir_xt = (1.0 - a_prev - sigma ** 2).sqrt() * e_t
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt + sigma * noise.to(dt) * temperature
    return x_prev, pred_x0


@torch.no_grad()
def ddim_sample(unet_fn, buf, x_T, cond, uncond, S=50, spacing="uniform_trailing", eta=1.0,
                cfg_scale=7.5, guidance_rescale=0.0, fs=None, noises=None, precision=None,
                record=None, uncond_img=None, cfg_img=None):
    """ddim.py:141-215 (and, with `uncond_img`, ddim_multiplecond.py:136-208: third forward per step).  `unet_fn(xc, t, context, fs)` is the denoiser;
    cond/uncond = dict(c_crossattn=[ctx], c_concat=[z]) (hybrid conditioning, ddpm3d.py:1077-1081).
    `noises[i]` is the N(0,1) draw of loop iteration i (the reference calls torch.randn each step)."""
    tab = ddim_tables(buf, S, spacing, eta)
    x = x_T if precision is None else x_T.to(precision)
    steps = np.flip(tab["timesteps"])
    b = x.shape[0]

    def apply(c):
        xc = torch.cat([x] + list(c["c_concat"]), dim=1)
        cc = torch.cat(list(c["c_crossattn"]), dim=1)
        t = torch.full((b,), int(step), dtype=torch.long, device=x.device)
        return unet_fn(xc, t, cc, fs)

    pred_x0 = x
    for i, step in enumerate(steps):
        index = S - i - 1
        v_c = apply(cond)
        v_u = apply(uncond) if (uncond is not None and cfg_scale != 1.0) else None
        v_ui = apply(uncond_img) if (v_u is not None and uncond_img is not None) else None
        noise = noises[i] if noises is not None else torch.randn(x.shape, device=x.device)
        if record is not None:
            record.append(dict(x=x, v_c=v_c, v_u=v_u))
        x, pred_x0 = ddim_step(buf, tab, x, v_c, v_u, step, index, noise, cfg_scale, guidance_rescale,
                               v_uncond_img=v_ui, cfg_img=cfg_img)
    return x, pred_x0
