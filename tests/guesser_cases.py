"""
The prior-sampling / guesser cases shared by oracle/gen_golden.py (which runs
them against the reference to write tests/golden/guessers.npz) and
tests/test_guessers.py (which runs them against ngmix_b200).
"""
import numpy as np


def guesser_cases(mod, rng_of):
    """the prior-sampling guesser cases, written once and run against the
    reference here and against ngmix_b200 in tests/test_guessers.py (`mod` is
    the package, `rng_of(seed)` a fresh RandomState)"""
    pri, jp, gs = mod.priors, mod.joint_prior, mod.guessers

    def joint(kind, rng, nband=1, tight=False):
        cen = pri.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng)
        gp = pri.GPriorBA(0.3, rng=rng)
        if tight:
            Tp = pri.FlatPrior(0.0, 0.3, rng=rng)
            Fp = pri.FlatPrior(0.0, 80.0, rng=rng)
        else:
            Tp = pri.TwoSidedErf(-1.0, 0.1, 1.0e3, 1.0, rng=rng)
            Fp = pri.TwoSidedErf(-1.0e2, 1.0, 1.0e9, 1.0e8, rng=rng)
        F = [Fp] * nband if nband > 1 else Fp
        if kind == "simple":
            return jp.PriorSimpleSep(cen, gp, Tp, F)
        fd = pri.Normal(0.5, 0.1, rng=rng)
        if kind == "bdf":
            return jp.PriorBDFSep(cen, gp, Tp, fd, F)
        return jp.PriorBDSep(cen, gp, Tp, pri.Normal(0.0, 0.1, rng=rng), fd, F)

    out = {}
    for kind in ("simple", "bdf", "bd"):
        rng = rng_of(100)
        p = joint(kind, rng)
        out["sample_%s_n5" % kind] = p.sample(5)
        out["sample_%s_one" % kind] = p.sample()
    rng = rng_of(101)
    out["erf_sample"] = pri.TwoSidedErf(-1.0, 0.5, 2.0, 0.25, rng=rng).sample(40)
    rng = rng_of(102)
    gp = pri.GPriorBA(0.2, rng=rng)
    g1, g2 = gp.sample2d(30)
    out["ba_sample2d"] = np.array([g1, g2])
    out["ba_sample2d_again"] = np.array(gp.sample2d(7))
    rng = rng_of(103)
    p = joint("simple", rng, nband=2)
    out["tflux_and_prior"] = gs.TFluxAndPriorGuesser(
        rng=rng, T=0.4, flux=[100.0, 60.0], prior=p)(nrand=4)
    rng = rng_of(104)
    out["tflux_fix"] = gs.TFluxGuesser(
        rng=rng, T=0.25, flux=100.0,
        prior=joint("simple", rng, tight=True))(nrand=6)
    rng = rng_of(105)
    out["tflux_and_prior_fix"] = gs.TFluxAndPriorGuesser(
        rng=rng, T=0.29, flux=75.0,
        prior=joint("simple", rng, tight=True))(nrand=6)
    rng = rng_of(106)
    out["bdf_guess"] = gs.BDFGuesser(T=0.4, flux=100.0,
                                     prior=joint("bdf", rng))(nrand=3)
    rng = rng_of(107)
    out["bd_guess"] = gs.BDGuesser(T=0.4, flux=100.0,
                                   prior=joint("bd", rng))(nrand=3)
    rng = rng_of(108)
    out["prior_guess"] = gs.PriorGuesser(joint("bdf", rng))(nrand=2)
    rng = rng_of(109)
    out["pars_guess_fix"] = gs.ParsGuesser(
        rng=rng, pars=[0.0, 0.0, 0.1, -0.2, 0.28, 78.0],
        prior=joint("simple", rng, tight=True))(nrand=8)
    return out
