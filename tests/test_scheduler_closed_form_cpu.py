"""Scheduler known answers derived INDEPENDENTLY of the oracle and of the product (pure Python ints, float64 math and
rationals from the published PNDM / DDIM formulas and the shipped scheduler_config.json values: beta_start 0.00085,
beta_end 0.012, scaled_linear, 1000 train steps, steps_offset 1, leading spacing, skip_prk_steps, set_alpha_to_one
false -- /root/reference/outputs/models/{denoising,inpainting}/best/scheduler/scheduler_config.json).  Both the oracle's
and the product's schedulers are checked against them, so the two cannot agree by sharing a mistake in these tables."""
import math
from fractions import Fraction

import pytest

N_TRAIN, B0, B1 = 1000, 0.00085, 0.012


def alpha_bar_f64():
    s0, s1 = math.sqrt(B0), math.sqrt(B1)
    out, acc = [], 1.0
    for i in range(N_TRAIN):
        beta = (s0 + (s1 - s0) * i / (N_TRAIN - 1)) ** 2
        acc *= 1.0 - beta
        out.append(acc)
    return out


def plms_list(n):
    r = N_TRAIN // n
    ts = [i * r + 1 for i in range(n)]
    return list(reversed(ts[:-1] + [ts[-2]] + [ts[-1]]))


def ddim_list(n):
    r = N_TRAIN // n
    return [i * r + 1 for i in reversed(range(n))]


def sliced(lst, n, strength):
    t_start = max(n - min(int(n * strength), n), 0)
    return lst[t_start:]


# literal expectations (written out, not computed)
DENOISE_FULL = [951, 901, 901, 851, 801, 751, 701, 651, 601, 551, 501, 451, 401, 351, 301, 251, 201, 151, 101, 51, 1]
DENOISE_RUN = [501, 451, 401, 351, 301, 251, 201, 151, 101, 51, 1]
COLORIZE_RUN = [727, 694, 661, 628, 595, 562, 529, 496, 463, 430, 397, 364, 331, 298, 265, 232, 199, 166, 133, 100, 67, 34, 1]
INPAINT_RUN = [562, 529, 496, 463, 430, 397, 364, 331, 298, 265, 232, 199, 166, 133, 100, 67, 34, 1]
SR20_RUN = [801, 751, 701, 651, 601, 551, 501, 451, 401, 351, 301, 251, 201, 151, 101, 51, 1]


def test_literal_timestep_lists_follow_from_the_formulas():
    assert plms_list(20) == DENOISE_FULL and sliced(plms_list(20), 20, 0.5) == DENOISE_RUN
    assert sliced(plms_list(30), 30, 0.75) == COLORIZE_RUN and len(COLORIZE_RUN) == 23
    assert sliced(ddim_list(30), 30, 0.6) == INPAINT_RUN and len(INPAINT_RUN) == 18
    assert sliced(plms_list(20), 20, 0.8) == SR20_RUN and len(SR20_RUN) == 17
    s50 = sliced(plms_list(50), 50, 0.8)
    assert len(s50) == 41 and s50[0] == 801 and s50[-1] == 1 and s50[1] == 781


def _both_scheduler_families():
    from oracle import schedulers as osch
    from image_restoration_and_enhancement_b200 import schedulers as psch
    return (("oracle", osch.PNDMScheduler, osch.DDIMScheduler), ("product", psch.PNDMScheduler, psch.DDIMScheduler))


@pytest.mark.parametrize("which", [0, 1])
def test_timesteps_and_alpha_bar_match_the_closed_form(which):
    name, PNDM, DDIM = _both_scheduler_families()[which]
    ab = alpha_bar_f64()
    p = PNDM()
    for i in (0, 1, 250, 500, 501, 951, 999):
        assert abs(float(p.alphas_cumprod[i]) - ab[i]) <= 3e-6 * ab[i] + 1e-9, (name, i)     # float32 table vs float64
    assert abs(ab[0] - (1 - B0)) < 1e-15 and 0.0046 < ab[999] < 0.0047
    for n, strength, want in ((20, 0.5, DENOISE_RUN), (30, 0.75, COLORIZE_RUN), (20, 0.8, SR20_RUN)):
        p.set_timesteps(n)
        full = [int(t) for t in p.timesteps]
        assert full == plms_list(n), name
        assert sliced(full, n, strength) == want
    d = DDIM()
    d.set_timesteps(30)
    assert [int(t) for t in d.timesteps] == ddim_list(30) and sliced(ddim_list(30), 30, 0.6) == INPAINT_RUN


def _plms_reference_run(timesteps, n_steps):
    """step_plms unrolled with rationals for the history weights and float64 for the alpha-bar coefficients: returns per
    call (UNet timestep, {history age: weight}, use_cur, c_sample, c_eps) -- prev = c_sample*sample - c_eps*e_mix."""
    ab = alpha_bar_f64()
    r = N_TRAIN // n_steps
    A = lambda t: ab[t] if t >= 0 else ab[0]            # set_alpha_to_one = false -> final_alpha_cumprod = alpha_bar[0]
    out, n_ets = [], 0
    for counter, t_in in enumerate(timesteps):
        t, prev_t = t_in, t_in - r
        if counter != 1:
            n_ets = min(n_ets + 1, 4)
        else:
            prev_t, t = t, t + r
        if n_ets == 1 and counter == 0:
            w, use_cur = {0: Fraction(1)}, False
        elif n_ets == 1 and counter == 1:
            w, use_cur = {0: Fraction(1, 2), 1: Fraction(1, 2)}, True            # (eps + ets[-1]) / 2 from cur_sample
        elif n_ets == 2:
            w, use_cur = {0: Fraction(3, 2), 1: Fraction(-1, 2)}, False
        elif n_ets == 3:
            w, use_cur = {0: Fraction(23, 12), 1: Fraction(-16, 12), 2: Fraction(5, 12)}, False
        else:
            w, use_cur = {0: Fraction(55, 24), 1: Fraction(-59, 24), 2: Fraction(37, 24), 3: Fraction(-9, 24)}, False
        a_t, a_p = A(t), A(prev_t)
        c_sample = math.sqrt(a_p / a_t)
        c_eps = (a_p - a_t) / (a_t * math.sqrt(1 - a_p) + math.sqrt(a_t * (1 - a_t) * a_p))
        out.append((t_in, w, use_cur, c_sample, c_eps))
    return out


def test_product_plms_plans_match_rational_weights_and_float64_coefficients():
    from image_restoration_and_enhancement_b200.schedulers import PNDMScheduler
    for n, strength in ((20, 0.5), (30, 0.75), (50, 0.8)):
        s = PNDMScheduler()
        s.set_timesteps(n)
        ts = s.get_timesteps(n, strength)
        plans = s.plan(ts)
        ref = _plms_reference_run(ts, n)
        assert len(plans) == len(ref)
        slots_by_age: list[int] = []                    # newest first: slot holding the eps of age 1, 2, 3
        for k, (p, (t_in, w, use_cur, cs, ce)) in enumerate(zip(plans, ref)):
            assert p.timestep == t_in and p.use_cur == use_cur and p.save_cur == (k == 0)
            got = {0: Fraction(p.w[4]).limit_denominator(48)}
            for age, slot in enumerate(slots_by_age[:3], start=1):
                if p.w[slot] != 0.0 and not (k != 1 and slot == p.store_slot):
                    got[age] = Fraction(p.w[slot]).limit_denominator(48)
            assert got == w, (n, k, got, w)
            # c_eps carries (a_p - a_t) of two float32 table entries (diffusers' arithmetic): ~7e-5 relative when they
            # differ by < 1e-3 (the last step), so 2e-4
            assert abs(p.c_sample - cs) <= 2e-6 * abs(cs) and abs(p.c_eps - ce) <= 2e-4 * abs(ce) + 1e-9, (n, k)
            if p.store_slot >= 0:
                slots_by_age.insert(0, p.store_slot)
        # the img2img quirk: call #2 re-does step 1 Heun-style, and the run ends on final_alpha_cumprod (prev_t < 0)
        assert ref[1][2] is True and plans[1].store_slot == -1


def test_product_ddim_plans_match_float64_coefficients():
    from image_restoration_and_enhancement_b200.schedulers import DDIMScheduler
    ab = alpha_bar_f64()
    s = DDIMScheduler()
    s.set_timesteps(30)
    ts = s.get_timesteps(30, 0.6)
    for p, t in zip(s.plan(ts), INPAINT_RUN):
        a_t = ab[t]
        a_p = ab[t - 33] if t - 33 >= 0 else ab[0]
        # prev = sqrt(a_p) (x - sqrt(1-a_t) e)/sqrt(a_t) + sqrt(1-a_p) e = c_sample x - c_eps e
        cs = math.sqrt(a_p / a_t)
        ce = math.sqrt(a_p) * math.sqrt(1 - a_t) / math.sqrt(a_t) - math.sqrt(1 - a_p)
        assert p.timestep == t and abs(p.c_sample - cs) <= 2e-6 * cs and abs(p.c_eps - ce) <= 2e-4 * abs(ce) + 1e-9
        assert p.w[4] == 1.0 and p.store_slot == -1


def test_oracle_schedulers_step_like_the_closed_form():
    """One numeric trajectory: the oracle's PNDM / DDIM ``step`` on scalar 'tensors' against the float64 recurrences."""
    import torch
    from oracle.schedulers import DDIMScheduler, PNDMScheduler, get_timesteps
    ab = alpha_bar_f64()
    eps_of = lambda k: math.sin(1.0 + 0.7 * k)                                  # any deterministic eps sequence
    # ---- PLMS, denoise config
    s = PNDMScheduler(); s.set_timesteps(20)
    ts, _ = get_timesteps(s, 20, 0.5)
    x = torch.tensor([0.3], dtype=torch.float32)
    ref = _plms_reference_run([int(t) for t in ts], 20)
    xr, hist, cur = 0.3, [], None
    for k, t in enumerate(ts):
        e = eps_of(k)
        x = s.step(torch.tensor([e], dtype=torch.float32), t, x)
        _, w, use_cur, cs, ce = ref[k]
        if k != 1:
            hist.insert(0, e)
            mix = sum(float(wt) * hist[age] for age, wt in w.items())
        else:
            mix = float(w[0]) * e + float(w[1]) * hist[0]
        base = cur if use_cur else xr
        if k == 0:
            cur = xr
        xr = cs * base - ce * mix
        assert abs(float(x) - xr) <= 2e-5 * max(1.0, abs(xr)), (k, float(x), xr)
    # ---- DDIM, inpaint config
    d = DDIMScheduler(); d.set_timesteps(30)
    ts, _ = get_timesteps(d, 30, 0.6)
    x, xr = torch.tensor([-0.2], dtype=torch.float32), -0.2
    for k, t in enumerate(ts):
        e, t = eps_of(k), int(t)
        x = d.step(torch.tensor([e], dtype=torch.float32), t, x)
        a_t, a_p = ab[t], (ab[t - 33] if t - 33 >= 0 else ab[0])
        x0 = (xr - math.sqrt(1 - a_t) * e) / math.sqrt(a_t)
        xr = math.sqrt(a_p) * x0 + math.sqrt(1 - a_p) * e
        assert abs(float(x) - xr) <= 2e-5 * max(1.0, abs(xr)), (k, float(x), xr)
