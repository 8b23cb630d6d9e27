"""Mirror of ``motion_blur/generate_trajectory.py``: Boracchi-Foi random camera-shake trajectories.

The trajectory is a 2000-step random walk whose every step consumes numpy's *global* MT19937 stream
(generate_trajectory.py:48-76), conditionally (the impulsive "big shake" draws one extra uniform), so it is
inherently sequential host work and the seeded samples are the INPUT of the GPU rasteriser
(psf_ops.rasterize_psfs), exactly as BASELINE.json's north_star lays out.  Same constructor, attributes and
RNG consumption as the reference class, so a seeded run yields bit-identical ``x``.
"""
import numpy as np


class Trajectory(object):
    def __init__(self, canvas=64, iters=2000, max_len=60, expl=None):
        self.canvas = canvas
        self.iters = iters
        self.max_len = max_len
        # generate_trajectory.py:27-30: an unspecified `expl` is itself a draw from the global stream
        self.expl = 0.1 * np.random.uniform(0, 1) if expl is None else expl
        self.tot_length = None
        self.big_expl_count = None
        self.x = None
        self.unprocessedX = None

    def fit(self):
        """Draw one trajectory (generate_trajectory.py:38-98).  Returns self; samples in ``x`` (complex128[iters])."""
        rnd = np.random
        iters, expl = self.iters, self.expl
        step = self.max_len / (iters - 1)
        # four shape parameters, in the reference's draw order (:46-53)
        centripetal = 0.7 * rnd.uniform(0, 1)
        prob_big_shake = 0.2 * rnd.uniform(0, 1)
        gaussian_shake = 10 * rnd.uniform(0, 1)
        init_angle = 360 * rnd.uniform(0, 1)
        ang = np.deg2rad(init_angle)
        v0 = complex(np.cos(ang), np.sin(ang))
        v = v0 * self.max_len / (iters - 1)
        if expl > 0:
            v = v0 * expl
        x = np.zeros(iters, dtype=np.complex128)
        tot_length = 0
        big = 0
        shake_threshold = prob_big_shake * expl
        norm_len = self.max_len / float(iters - 1)
        for t in range(iters - 1):
            if rnd.uniform() < shake_threshold:
                # impulsive perturbation: roughly reverse the velocity (:70-72)
                kick = 2 * v * (np.exp(complex(0, np.pi + (rnd.uniform() - 0.5))))
                big += 1
            else:
                kick = 0
            g = complex(rnd.randn(), rnd.randn())
            # x[t] is a numpy complex128 scalar, which makes dv and v numpy scalars too: numpy divides a complex by
            # multiplying with the reciprocal, CPython by dividing, and the reference's bits come from the former
            dv = kick + expl * (gaussian_shake * g - centripetal * x[t]) * step
            v += dv
            v = (v / float(np.abs(v))) * norm_len
            x[t + 1] = x[t] + v
            tot_length = tot_length + abs(x[t + 1] - x[t])
        self.unprocessedX = np.copy(x)
        self.x = x + complex(self.canvas / 2, self.canvas / 2)      # start point at the canvas centre (:92)
        self.tot_length = tot_length
        self.big_expl_count = big
        return self

    def applyscale_factor(self):
        """generate_trajectory.py:100-104: rescale the raw walk so it just fits the canvas."""
        x = self.unprocessedX
        half = (self.canvas / 2) - 2
        s = np.max([np.max(-1 * x.real / half), np.max(-1 * x.imag / half), np.max(x.real / half), np.max(x.imag / half)])
        self.x = x / s + complex(self.canvas / 2, self.canvas / 2)
