"""Hand-derived known-answer cases (none exist upstream).  Each case is a tiny network with explicit
weights and an input whose outputs follow from the reference's formulas by inspection -- no oracle, no reference run.
Used twice: tests/test_oracle_known_answers.py pins the oracle (CPU), tests/test_gpu_known_answers.py pins the CUDA path.
"""
import numpy as np


def impulse_conv():
    """One impulse at x[2,3,c=1] through a valid 3x3 conv (cross-correlation, filters (F,k,k,C), Classes/CNNModel.py:227-240):
    y[i,j,f] = w[f, 2-i, 3-j, 1] + b[f] where that tap exists, else b[f].  All values positive => LeakyReLU is the identity."""
    x = np.zeros((1, 6, 6, 2), np.float32)
    x[0, 2, 3, 1] = 1.0
    w = (np.arange(2 * 3 * 3 * 2, dtype=np.float64).reshape(2, 3, 3, 2) + 1.0) / 10.0          # all distinct, > 0
    b = np.array([0.5, 0.25])
    want = np.zeros((1, 4, 4, 2))
    for f in range(2):
        want[0, :, :, f] = b[f]
        for i in range(4):
            for j in range(4):
                u, v = 2 - i, 3 - j
                if 0 <= u < 3 and 0 <= v < 3:
                    want[0, i, j, f] += w[f, u, v, 1]
    return dict(input_shape=(6, 6, 2), conv=[(2, 3)], hidden=[], conv_w=[w], conv_b=[b],
                dense_w=[np.zeros((2, 2 * 2 * 2))], dense_b=[np.zeros(2)], x=x, conv_out=want)


def pool_tie():
    """2x2 window [[3,3],[1,3]] (1x1 identity conv): pooled 3; the gradient of logit 0 w.r.t. the conv output is
    W_out[0, window] at EVERY maximum under the NumPy rule (switches = patch == max, Classes/CNNModel.py:260,274-275) and at the
    FIRST maximum (row-major) under the torch rule."""
    x = np.zeros((1, 2, 4, 1), np.float32)
    x[0, :, 0:2, 0] = [[3, 3], [1, 3]]
    x[0, :, 2:4, 0] = [[2, 5], [5, 4]]                      # second window: tie between (0,1) and (1,0)
    w_out = np.array([[0.7, -0.4], [0.1, 0.2]])             # logits = W_out . [pool0, pool1]
    dup = np.zeros((1, 2, 4, 1))
    dup[0, :, 0:2, 0] = [[0.7, 0.7], [0.0, 0.7]]
    dup[0, :, 2:4, 0] = [[0.0, -0.4], [-0.4, 0.0]]
    first = np.zeros((1, 2, 4, 1))
    first[0, 0, 0, 0] = 0.7
    first[0, 0, 3, 0] = -0.4
    return dict(input_shape=(2, 4, 1), conv=[(1, 1)], hidden=[], conv_w=[np.ones((1, 1, 1, 1))], conv_b=[np.zeros(1)],
                dense_w=[w_out], dense_b=[np.zeros(2)], x=x, pooled=np.array([[3.0, 5.0]]), dA_dup=dup, dA_first=first)


def flatten_order():
    """Pooled map (2,2,2) with value 10*y + x + 100*c; a one-hot dense row picks flat index j: HWC index = (y*2+x)*2+c
    (Classes/CNNModel.py:178), CHW index = c*4 + y*2 + x (ADCNNM.py:77)."""
    x = np.zeros((1, 4, 4, 2), np.float32)
    for y in range(2):
        for xx in range(2):
            for c in range(2):
                x[0, 2 * y:2 * y + 2, 2 * xx:2 * xx + 2, c] = 1.0 + 10 * y + xx + 100 * c      # constant windows: pool = value
    j = 5
    w = np.zeros((2, 8))
    w[0, j] = 1.0
    hwc = 1.0 + 10 * ((j // 2) // 2) + ((j // 2) % 2) + 100 * (j % 2)          # j=5 -> pixel 2 (y=1,x=0), c=1 -> 111
    chw = 1.0 + 10 * ((j % 4) // 2) + ((j % 4) % 2) + 100 * (j // 4)           # j=5 -> c=1, pixel 1 (y=0,x=1) -> 102
    eye = np.zeros((2, 1, 1, 2))
    eye[0, 0, 0, 0] = eye[1, 0, 0, 1] = 1.0
    return dict(input_shape=(4, 4, 2), conv=[(2, 1)], hidden=[], conv_w=[eye], conv_b=[np.zeros(2)], dense_w=[w],
                dense_b=[np.zeros(2)], x=x, logit0_hwc=hwc, logit0_chw=chw)


def softmax_clip():
    """Logits (60, 0): the NumPy head clips to +-50 first (Classes/CNNModel.py:203-212): p1 = e^-50 / (1 + e^-50 + 1e-12);
    a plain softmax gives e^-60 / (1 + e^-60)."""
    x = np.ones((1, 2, 2, 1), np.float32)
    clipped = np.exp(-50.0) / (1.0 + np.exp(-50.0) + 1e-12)
    plain = np.exp(-60.0) / (1.0 + np.exp(-60.0))
    return dict(input_shape=(2, 2, 1), conv=[(1, 1)], hidden=[], conv_w=[np.ones((1, 1, 1, 1))], conv_b=[np.zeros(1)],
                dense_w=[np.zeros((2, 1))], dense_b=[np.array([60.0, 0.0])], x=x, p1_clip=clipped, p1_plain=plain)


def leaky_at_zero(alpha=0.25):
    """A hidden unit whose pre-activation is EXACTLY 0: the reference tests z > 0 (strict, Classes/CNNModel.py:184,
    explainability.py:28-29), so the activation is alpha*0 = 0 and its derivative is alpha.  Pooled = [1, 1], hidden w = [1, -1]
    => z = 0; logits = W_out * h + b.  d(logit_0)/d(pooled) = w_hidden^T * (W_out[0] * alpha) = [alpha*2, -alpha*2]."""
    x = np.ones((1, 2, 4, 1), np.float32)
    return dict(input_shape=(2, 4, 1), conv=[(1, 1)], hidden=[1], conv_w=[np.ones((1, 1, 1, 1))], conv_b=[np.zeros(1)],
                dense_w=[np.array([[1.0, -1.0]]), np.array([[2.0], [-1.0]])], dense_b=[np.zeros(1), np.array([0.3, 0.1])], x=x,
                alpha=alpha, logits=np.array([[0.3, 0.1]]), g_pool=np.array([2.0 * alpha, -2.0 * alpha]))
