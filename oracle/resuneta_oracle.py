"""CPU oracle for the ResUnet-a multitask hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* of the algorithm of thimabru1010/ResUnet-a_mltsk_keras
on plain torch-CPU tensors (fp32 or fp64).  It is the checker for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under the product package
``resunet-a_mltsk_keras_b200/`` imports this module.

PARITY UNPINNED at the TensorFlow boundary: TensorFlow/Keras is not installable in
the build container and the reference ships no test or golden vector for network,
loss or optimizer numerics.  What *is* pinned:
  * metrics-from-confusion-matrix against the four matrices printed in
    ``infos_training_train_on_batch.txt`` (tests/golden/kat_metrics.json),
  * the conv/BN/pool conventions against a pure-numpy direct-loop restatement
    (``oracle/numpy_loops.py``),
  * analytic known answers T1..T4, G1..G5 (tests/test_oracle_cpu.py).

Reference lines followed (all relative to /root/reference):
  ResUnet_a/model2.py:14-193   primary graph ("v2")
  ResUnet_a/model.py:14-171    older graph   ("v1")
  multitasking_utils.py:38-85  Tanimoto loss / dual
  utils.py:466-491             weighted categorical cross-entropy
  train_ISPRS.py:404-452       optimizer / loss / metric wiring
  test_ISPRS.py:26-36,48-87,102-152,295-314 ; utils.py:52-57   inference side
Keras layer defaults that are not visible in the reference source are listed in
SURVEY.md §A.2 and hard-coded here (BN eps 1e-3 / momentum 0.99, Glorot-uniform, ...).

Layout: the public functions take and return NHWC tensors like Keras; torch's NCHW is
used internally.  Parameters live in an ordered dict keyed by the Keras auto-names
(``conv2d_7/kernel`` HWIO, ``batch_normalization_3/gamma`` ...).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3        # keras BatchNormalization default epsilon
BN_MOMENTUM = 0.99   # keras BatchNormalization default momentum

ENC_FILTERS = [32, 64, 128, 256, 512, 1024]
ENC_DILATIONS = [[1, 3, 15, 31], [1, 3, 15, 31], [1, 3, 15], [1, 3, 15], [1], [1]]


# --------------------------------------------------------------------------------------
# graph walker: one routine, run either to create parameters (in Keras creation order) or
# to evaluate the network.  Mirrors build_model_ResUneta (model2.py:14-193 / model.py:14-171)
# --------------------------------------------------------------------------------------
class _Namer:
    """Keras auto-naming: conv2d, conv2d_1, ... / batch_normalization, batch_normalization_1, ..."""

    def __init__(self):
        self.n = {}

    def __call__(self, base):
        i = self.n.get(base, 0)
        self.n[base] = i + 1
        return base if i == 0 else f"{base}_{i}"


class _Ctx:
    def __init__(self, params, training, create, dtype, gen, state_out):
        self.p = params
        self.training = training
        self.create = create
        self.dtype = dtype
        self.gen = gen
        self.namer = _Namer()
        self.state_out = state_out  # new moving stats (training mode)

    # -- layers ------------------------------------------------------------------------
    def conv(self, x, cout, k=1, stride=1, dilation=1, padding="valid", name=None):
        """keras Conv2D, use_bias=True, kernel HWIO, Glorot-uniform init."""
        name = name or self.namer("conv2d")
        cin = x.shape[1]
        if self.create:
            limit = math.sqrt(6.0 / (k * k * cin + k * k * cout))
            w = (torch.rand(k, k, cin, cout, generator=self.gen, dtype=torch.float64) * 2 - 1) * limit
            self.p[name + "/kernel"] = w.to(self.dtype)
            self.p[name + "/bias"] = torch.zeros(cout, dtype=self.dtype)
        w = self.p[name + "/kernel"].permute(3, 2, 0, 1)  # HWIO -> OIHW
        b = self.p[name + "/bias"]
        pad = 0
        if padding == "same":
            # stride-1 'same' with kernel 3, dilation d -> symmetric zero pad of d
            assert stride == 1
            pad = dilation * (k - 1) // 2
        if self.create:  # shape walk only
            ho = (x.shape[2] + 2 * pad - dilation * (k - 1) - 1) // stride + 1
            wo = (x.shape[3] + 2 * pad - dilation * (k - 1) - 1) // stride + 1
            return torch.zeros(x.shape[0], cout, ho, wo, dtype=self.dtype)
        return F.conv2d(x, w, b, stride=stride, padding=pad, dilation=dilation)

    def bn(self, x):
        """keras BatchNormalization(axis=-1, momentum=.99, epsilon=1e-3)."""
        name = self.namer("batch_normalization")
        c = x.shape[1]
        if self.create:
            self.p[name + "/gamma"] = torch.ones(c, dtype=self.dtype)
            self.p[name + "/beta"] = torch.zeros(c, dtype=self.dtype)
            self.p[name + "/moving_mean"] = torch.zeros(c, dtype=self.dtype)
            self.p[name + "/moving_variance"] = torch.ones(c, dtype=self.dtype)
        g = self.p[name + "/gamma"].view(1, c, 1, 1)
        be = self.p[name + "/beta"].view(1, c, 1, 1)
        if self.training:
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
            n = x.numel() // c
            if self.state_out is not None:
                with torch.no_grad():
                    # TF FusedBatchNorm feeds the Bessel-corrected variance to the moving average
                    var_u = var * (n / max(n - 1, 1))
                    mm = self.p[name + "/moving_mean"]
                    mv = self.p[name + "/moving_variance"]
                    self.state_out[name + "/moving_mean"] = (mm * BN_MOMENTUM + mean * (1 - BN_MOMENTUM)).detach()
                    self.state_out[name + "/moving_variance"] = (mv * BN_MOMENTUM + var_u * (1 - BN_MOMENTUM)).detach()
        else:
            mean = self.p[name + "/moving_mean"]
            var = self.p[name + "/moving_variance"]
        inv = torch.rsqrt(var + BN_EPS).view(1, c, 1, 1)
        return (x - mean.view(1, c, 1, 1)) * inv * g + be


def _up(x, k):
    return x if k == 1 else F.interpolate(x, scale_factor=k, mode="nearest")


def _pool(x, k):
    return x if k == 1 else F.max_pool2d(x, k, k)


def _resblock(cx, x, f, dils, identity):
    """model2.py:15-34 (identity add) / model.py:15-33 (no identity; single branch passes through)."""
    outs = [x] if identity else []
    for d in dils:
        t = cx.bn(x)
        t = F.relu(t)
        t = cx.conv(t, f, 3, 1, d, "same")
        t = cx.bn(t)
        t = F.relu(t)
        t = cx.conv(t, f, 3, 1, d, "same")
        outs.append(t)
    out = outs[0]
    for t in outs[1:]:
        out = out + t
    return out


def _psp_levels(img_width):
    # gated on the *input patch* width, model2.py:49-52
    lv = [1, 2]
    if img_width >= 128:
        lv.append(4)
    if img_width >= 256:
        lv.append(8)
    return lv


def _psp_v2(cx, x, f, img_width):
    """model2.py:41-79: pool -> up -> Conv2DN(f/4) per level; concat with x; Conv2DN(f)."""
    ups = [_up(_pool(x, k), k) for k in _psp_levels(img_width)]
    br = []
    for u in ups:
        t = cx.conv(u, f // 4)
        t = cx.bn(t)
        br.append(t)
    t = torch.cat(br + [x], dim=1)
    t = cx.conv(t, f)
    return cx.bn(t)


def _psp_v1(cx, x, f, img_width):
    """model.py:35-64: pool -> Conv(f/4) -> up per level; concat with x; Conv(f). No BN."""
    lv = _psp_levels(img_width)
    pooled = [_pool(x, k) for k in lv]
    convd = [cx.conv(p, f // 4) for p in pooled]
    ups = [_up(c, k) for c, k in zip(convd, lv)]
    t = torch.cat(ups + [x], dim=1)
    return cx.conv(t, f)


def _heads(cx, x_psp, x_comb, n, multitask):
    """model2.py:144-191 == model.py:116-169. Heads are independent."""
    if not multitask:
        t = cx.conv(x_psp, n)
        return F.softmax(t, dim=1)
    # creation order: seg1, seg2, seg3, bound conv x2, dist conv x3, color
    s = F.relu(cx.conv(x_psp, 32, 3, padding="same", name="seg1"))   # ZeroPadding2D(1)+valid == same
    s = F.relu(cx.conv(s, 32, 3, padding="same", name="seg2"))
    s = cx.conv(s, n, name="seg3")
    seg = F.softmax(s, dim=1)
    b = F.relu(cx.conv(x_psp, 32, 3, padding="same"))
    b = cx.conv(b, n)
    bound = torch.sigmoid(b)
    d = F.relu(cx.conv(x_comb, 32, 3, padding="same"))
    d = F.relu(cx.conv(d, 32, 3, padding="same"))
    d = cx.conv(d, n)
    dist = F.softmax(d, dim=1)   # softmax, model2.py:182
    color = torch.sigmoid(cx.conv(x_comb, 3, name="color"))
    return OrderedDict(seg=seg, bound=bound, dist=dist, color=color)


def _network(cx, x, variant, num_classes, multitask, img_width):
    v2 = variant == "v2"
    skips = []
    t = cx.conv(x, 32)                                   # stem, model2.py:101
    skips.append(t)                                      # c1
    for lvl, (f, dils) in enumerate(zip(ENC_FILTERS, ENC_DILATIONS)):
        if lvl > 0:
            t = cx.conv(t, f, 1, stride=2)               # model2.py:103-111 (no BN, no act)
        t = _resblock(cx, t, f, dils, identity=v2)
        if lvl < 5:
            skips.append(t)                              # c2..c6
    if v2:
        t = F.relu(_psp_v2(cx, t, 1024, img_width))      # model2.py:114-116
    else:
        t = _psp_v1(cx, t, 1024, img_width)              # model.py:90
    for lvl in range(4, -1, -1):                         # decoder f = 512,256,128,64,32
        f, dils = ENC_FILTERS[lvl], ENC_DILATIONS[lvl]
        skip = skips[lvl + 1]
        if v2:
            t = _up(t, 2)                                # model2.py:89-94
            t = cx.conv(t, f // 2)
            t = cx.bn(t)
            t = torch.cat([F.relu(t), skip], dim=1)      # combine, model2.py:81-87
            t = cx.conv(t, f)
            t = cx.bn(t)
        else:
            t = cx.conv(t, f)                            # model.py:93-94
            t = _up(t, 2)
            t = torch.cat([F.relu(t), skip], dim=1)      # model.py:66-70
            t = cx.conv(t, f)
        t = _resblock(cx, t, f, dils, identity=v2)
    t = torch.cat([F.relu(t), skips[0]], dim=1)          # x_comb = combine(x, c1, 32)
    t = cx.conv(t, 32)
    if v2:
        x_comb = cx.bn(t)
        x_psp = F.relu(_psp_v2(cx, x_comb, 32, img_width))
    else:
        x_comb = t
        x_psp = _psp_v1(cx, x_comb, 32, img_width)
    return _heads(cx, x_psp, x_comb, num_classes, multitask)


# --------------------------------------------------------------------------------------
# public model API
# --------------------------------------------------------------------------------------
def init_params(input_shape, num_classes, multitask=True, variant="v2", seed=1234, dtype=torch.float32):
    """Create parameters in Keras creation order with Keras default initialisers."""
    h, w, c = input_shape
    gen = torch.Generator().manual_seed(seed)
    params = OrderedDict()
    cx = _Ctx(params, training=True, create=True, dtype=dtype, gen=gen, state_out=None)
    with torch.no_grad():
        _network(cx, torch.zeros(1, c, h, w, dtype=dtype), variant, num_classes, multitask, w)
    return params


def is_trainable(name):
    return not (name.endswith("/moving_mean") or name.endswith("/moving_variance"))


def forward(params, x_nhwc, training, num_classes, multitask=True, variant="v2", new_state=None):
    """Forward pass.  x NHWC -> dict of NHWC tensors (multitask) or one NHWC tensor.

    ``training=True`` is what ``train_on_batch`` does (batch statistics); ``False`` is
    ``test_on_batch`` / ``predict`` (moving statistics).  If ``new_state`` is a dict it
    receives the updated moving statistics.
    """
    x = x_nhwc.permute(0, 3, 1, 2)
    cx = _Ctx(params, training=training, create=False, dtype=x.dtype, gen=None, state_out=new_state)
    out = _network(cx, x, variant, num_classes, multitask, x_nhwc.shape[2])
    if isinstance(out, dict):
        return OrderedDict((k, v.permute(0, 2, 3, 1)) for k, v in out.items())
    return out.permute(0, 2, 3, 1)


# --------------------------------------------------------------------------------------
# losses (NHWC, probabilities in, per Keras)
# --------------------------------------------------------------------------------------
def tanimoto_loss(label, pred):
    """multitasking_utils.py:38-68 — note the weights come from the FIRST argument."""
    smooth = 1e-5
    vli = label.sum(dim=(1, 2)).mean(dim=0)                    # :46
    wli = 1.0 / (vli ** 2)                                     # :47
    isinf = torch.isinf(wli)
    new_w = torch.where(isinf, torch.zeros_like(wli), wli)     # :52
    wli = torch.where(isinf, torch.ones_like(wli) * new_w.max(), wli)  # :53
    sum_square = (pred ** 2 + label ** 2).sum(dim=(1, 2))      # :56-59
    sum_product = (pred * label).sum(dim=(1, 2))               # :61-62
    num = (wli * sum_product).sum(dim=-1)                      # :63
    den = (wli * (sum_square - sum_product)).sum(dim=-1)       # :65-66
    return (num + smooth) / (den + smooth)                     # :67


def tanimoto_dual_loss(label, pred):
    """multitasking_utils.py:78-84.  First call has its arguments swapped (:79)."""
    l1 = tanimoto_loss(pred, label)
    l2 = tanimoto_loss(1.0 - label, 1.0 - pred)
    return 1.0 - 0.5 * (l1 + l2)            # shape [B]


def weighted_categorical_crossentropy(weights):
    """utils.py:466-491 -> fn(y_true, y_pred) -> [B,H,W]."""
    def loss(y_true, y_pred):
        w = torch.as_tensor(weights, dtype=y_pred.dtype)
        p = y_pred / y_pred.sum(dim=-1, keepdim=True)
        p = p.clamp(1e-7, 1 - 1e-7)
        return -(y_true * torch.log(p) * w).sum(dim=-1)
    return loss


def categorical_crossentropy(y_true, y_pred):
    p = y_pred / y_pred.sum(dim=-1, keepdim=True)
    p = p.clamp(1e-7, 1 - 1e-7)
    return -(y_true * torch.log(p)).sum(dim=-1)


def binary_crossentropy(y_true, y_pred):
    """keras BinaryCrossentropy: clip 1e-7, mean over last axis -> [B,H,W]."""
    p = y_pred.clamp(1e-7, 1 - 1e-7)
    return -(y_true * torch.log(p) + (1 - y_true) * torch.log(1 - p)).mean(dim=-1)


def mean_squared_error(y_true, y_pred):
    return ((y_pred - y_true) ** 2).mean(dim=-1)


LOSSES = {
    "tanimoto": tanimoto_dual_loss,
    "cce": categorical_crossentropy,
    "bce": binary_crossentropy,
    "mse": mean_squared_error,
}


def reduce_loss(per_elem):
    """keras SUM_OVER_BATCH_SIZE: mean over everything the loss fn returned."""
    return per_elem.mean()


def seg_metrics(y_true, y_pred):
    """train_ISPRS.py:446-449: categorical accuracy + TP/FP/TN/FN at threshold 0.5."""
    acc = (y_true.argmax(-1) == y_pred.argmax(-1)).to(torch.float64).mean().item()
    t = y_true > 0.5
    p = y_pred > 0.5
    tp = (t & p).sum().item()
    fp = (~t & p).sum().item()
    tn = (~t & ~p).sum().item()
    fn = (t & ~p).sum().item()
    return [acc, float(tp), float(fp), float(tn), float(fn)]


def total_loss(outputs, y, losses, loss_weights):
    """Σ_i w_i * mean(loss_i) — train_ISPRS.py:437-452. Returns (total, [per-head])."""
    per = []
    tot = 0.0
    for k in outputs:
        l = reduce_loss(losses[k](y[k], outputs[k]))
        per.append(l)
        tot = tot + loss_weights.get(k, 1.0) * l
    return tot, per


# --------------------------------------------------------------------------------------
# optimizers (Keras TF-2.2 forms)
# --------------------------------------------------------------------------------------
class Adam:
    def __init__(self, lr=1e-3, beta_1=0.9, beta_2=0.999, eps=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, eps
        self.t = 0
        self.m, self.v = {}, {}

    def step(self, params, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for k, g in grads.items():
            m = self.m.get(k, torch.zeros_like(g))
            v = self.v.get(k, torch.zeros_like(g))
            m = self.b1 * m + (1 - self.b1) * g
            v = self.b2 * v + (1 - self.b2) * g * g
            self.m[k], self.v[k] = m, v
            params[k] = (params[k] - lr_t * m / (v.sqrt() + self.eps)).detach()


class SGD:
    def __init__(self, lr=1e-3, momentum=0.8):
        self.lr, self.mom = lr, momentum
        self.v = {}

    def step(self, params, grads):
        for k, g in grads.items():
            v = self.v.get(k, torch.zeros_like(g))
            v = self.mom * v - self.lr * g
            self.v[k] = v
            params[k] = (params[k] + v).detach()


def loss_and_grads(params, x, y, losses, loss_weights, num_classes, multitask=True, variant="v2",
                   training=True):
    """One fwd+bwd. Returns (total, per_head, outputs, grads dict, new moving stats)."""
    leaf = OrderedDict()
    for k, v in params.items():
        leaf[k] = v.detach().clone().requires_grad_(is_trainable(k))
    new_state = {}
    out = forward(leaf, x, training, num_classes, multitask, variant, new_state)
    if not isinstance(out, dict):
        out = OrderedDict(seg=out)
        y = y if isinstance(y, dict) else {"seg": y}
        losses = losses if isinstance(losses, dict) else {"seg": losses}
    tot, per = total_loss(out, y, losses, loss_weights)
    names = [k for k in leaf if is_trainable(k)]
    gs = torch.autograd.grad(tot, [leaf[k] for k in names], allow_unused=True)
    grads = OrderedDict()
    for k, g in zip(names, gs):
        grads[k] = torch.zeros_like(leaf[k]) if g is None else g.detach()
    out = OrderedDict((k, v.detach()) for k, v in out.items())
    return tot.detach(), [p.detach() for p in per], out, grads, new_state


def train_on_batch(params, opt, x, y, losses, loss_weights, num_classes, multitask=True, variant="v2"):
    """train_ISPRS.py:148 semantics.  Mutates ``params`` in place, returns the Keras list
    [loss, seg_loss, bound_loss, dist_loss, color_loss, seg_acc, TP, FP, TN, FN] (:493-496)."""
    tot, per, out, grads, new_state = loss_and_grads(params, x, y, losses, loss_weights, num_classes,
                                                     multitask, variant, True)
    opt.step(params, grads)
    params.update(new_state)
    ydict = y if isinstance(y, dict) else {"seg": y}
    res = [tot.item()] + ([p.item() for p in per] if multitask else [])
    return res + seg_metrics(ydict["seg"], out["seg"])


def test_on_batch(params, x, y, losses, loss_weights, num_classes, multitask=True, variant="v2"):
    with torch.no_grad():
        out = forward(params, x, False, num_classes, multitask, variant)
        if not isinstance(out, dict):
            out = OrderedDict(seg=out)
            y = {"seg": y}
            losses = losses if isinstance(losses, dict) else {"seg": losses}
        tot, per = total_loss(out, y, losses, loss_weights)
    res = [tot.item()] + ([p.item() for p in per] if multitask else [])
    return res + seg_metrics(y["seg"], out["seg"])


# --------------------------------------------------------------------------------------
# inference side: chop / reconstruct / confusion / metrics
# --------------------------------------------------------------------------------------
def extract_patches(img, ps):
    """test_ISPRS.py:102-152: non-overlapping, stride = patch size, row-major, remainder dropped."""
    h, w = img.shape[:2]
    nh, nw = h // ps, w // ps
    out = np.zeros((nh * nw, ps, ps) + img.shape[2:], dtype=np.float64)
    c = 0
    for i in range(nh):
        for j in range(nw):
            out[c] = img[i * ps:(i + 1) * ps, j * ps:(j + 1) * ps]
            c += 1
    return out


def pred_reconstruction(ps, pred_labels, ref_shape):
    """test_ISPRS.py:48-68 (img_type=1): paste back row-major into zeros((H,W)) float64."""
    h, w = ref_shape
    nh, nw = h // ps, w // ps
    img = np.zeros((h, w))
    c = 0
    for i in range(nh):
        for j in range(nw):
            img[i * ps:(i + 1) * ps, j * ps:(j + 1) * ps] = pred_labels[c]
            c += 1
    return img


def confusion_matrix(y_true, y_pred, labels=None):
    """sklearn.metrics.confusion_matrix semantics (test_ISPRS.py:314): int64 counts, rows =
    true, cols = predicted, labels = sorted union of the values present unless given."""
    y_true = np.asarray(y_true).ravel()
    y_pred = np.asarray(y_pred).ravel()
    if labels is None:
        labels = np.union1d(np.unique(y_true), np.unique(y_pred))
    labels = np.asarray(labels)
    k = len(labels)
    cm = np.zeros((k, k), dtype=np.int64)
    t_idx = np.searchsorted(labels, y_true)
    p_idx = np.searchsorted(labels, y_pred)
    ok = (t_idx < k) & (p_idx < k)
    ok &= (labels[np.minimum(t_idx, k - 1)] == y_true) & (labels[np.minimum(p_idx, k - 1)] == y_pred)
    np.add.at(cm, (t_idx[ok], p_idx[ok]), 1)
    return cm


def metrics_from_confusion(cm):
    """utils.py:52-57 (accuracy_score, f1/recall/precision with average=None, all x100)
    expressed on the confusion matrix (rows = true)."""
    cm = np.asarray(cm, dtype=np.float64)
    diag = np.diag(cm)
    acc = 100.0 * diag.sum() / cm.sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        recall = np.where(cm.sum(1) > 0, diag / cm.sum(1), 0.0)
        precision = np.where(cm.sum(0) > 0, diag / cm.sum(0), 0.0)
        f1 = np.where(recall + precision > 0, 2 * recall * precision / (recall + precision), 0.0)
    return acc, 100 * f1, 100 * recall, 100 * precision


def compute_mcc(tp, tn, fp, fn):
    """train_ISPRS.py:30-32."""
    return (tp * tn - fp * fn) / math.sqrt((tp + fp) * (tp + fn) * (tn + fp) * (tn + fn))


# --------------------------------------------------------------------------------------
# synthetic workloads — shared by tests and bench so both sides see the same bytes
# --------------------------------------------------------------------------------------
def synth_batch(batch, hw, cin, num_classes, seed, block=16):
    """cfg2-style synthetic batch: x U[0,1); seg one-hot of randint smoothed into blocks;
    bound Bernoulli(.1); dist U[0,1); color U[0,1) 3ch.  Returns numpy float32 NHWC."""
    rng = np.random.RandomState(seed)
    x = rng.rand(batch, hw, hw, cin).astype(np.float32)
    nb = max(hw // block, 1)
    cls = rng.randint(0, num_classes, size=(batch, nb, nb))
    cls = np.repeat(np.repeat(cls, hw // nb, axis=1), hw // nb, axis=2)
    seg = np.eye(num_classes, dtype=np.float32)[cls]
    bound = (rng.rand(batch, hw, hw, num_classes) < 0.1).astype(np.float32)
    dist = rng.rand(batch, hw, hw, num_classes).astype(np.float32)
    color = rng.rand(batch, hw, hw, 3).astype(np.float32)
    return x, dict(seg=seg, bound=bound, dist=dist, color=color)
