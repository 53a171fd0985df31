"""MIMO oracle for the tests: OUT x IN reference FFTConvolvers (CPU oracle), outputs summed over
`in` in ascending order in f32 — the semantics SURVEY.md §8(e) defines for BASELINE configs[4]."""
import numpy as np

import oracle


class MimoOracle:
    def __init__(self, responses, block, max_len):
        self.n_out, self.n_in, _ = responses.shape
        self.conv = [[oracle.FFTConvolver.init(responses[o, i], block, max_len) for i in range(self.n_in)]
                     for o in range(self.n_out)]

    def process(self, x):  # x [IN][n] -> [OUT][n]
        n = x.shape[1]
        y = np.zeros((self.n_out, n), np.float32)
        tmp = np.zeros(n, np.float32)
        for o in range(self.n_out):
            for i in range(self.n_in):
                self.conv[o][i].process(x[i], tmp)
                y[o] = (y[o] + tmp).astype(np.float32)
        return y
