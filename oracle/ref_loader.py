"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference LTM modules.

Imports the unmodified reference files from ``/root/reference`` (read-only, only
present in the dev container) so that the clean-room oracle (``ltm_oracle.py``)
and the golden fixtures under ``tests/golden/`` can be pinned against them.
Nothing in the product package may import this module.

Scaffolding needed to import the files in isolation:
  * ``matplotlib`` is imported at module top of both LTM files
    (long_term_attention_gibbs.py:21, long_term_attention.py:21) but never used on
    the path and is not installed -> a stub is placed in ``sys.modules``.
  * both files use the relative import ``.basis_functions`` -> they are loaded
    under a synthetic package so that ``InfVideoLLaMA/__init__`` (which pulls
    omegaconf/decord/...) is never executed.
  * the Gaussian file references the undefined name ``ContinuousSoftmax``
    (long_term_attention.py:99,332; it lives in the un-vendored
    deep-spin/infinite-former repo).  A shim with the published forward
    (theta -> (mu, sigma^2) -> psi.integrate_psi_gaussian) is injected.
"""
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference exists only in the dev container; `oracle/install_ref.py` puts verbatim copies of the path's files
# under baseline/_ref/ (git-ignored, travels with the gpurun snapshot) so the same loader works on the GPU box
_CANDIDATES = [os.environ.get("INFLTM_REFERENCE_ROOT"), "/root/reference",
               os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
REF_ROOT = next((c for c in _CANDIDATES if c and os.path.isfile(os.path.join(
    c, "infty-Video-LLaMA", "InfVideoLLaMA", "models", "long_term_attention_gibbs.py"))), "/root/reference")
_VL = os.path.join(REF_ROOT, "infty-Video-LLaMA", "InfVideoLLaMA", "models")
_VC = os.path.join(REF_ROOT, "infty-VideoChat2", "models", "blip2")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_VL, "long_term_attention_gibbs.py"))


def _stub_matplotlib():
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        p = types.ModuleType("matplotlib.pyplot")
        m.pyplot = p
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = p


def _load_pkg(pkg_name: str, directory: str, files):
    _stub_matplotlib()
    if pkg_name not in sys.modules:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [directory]
        sys.modules[pkg_name] = pkg
    mods = {}
    for f in files:
        full = f"{pkg_name}.{f}"
        if full in sys.modules:
            mods[f] = sys.modules[full]
            continue
        spec = importlib.util.spec_from_file_location(full, os.path.join(directory, f + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        mods[f] = mod
    return mods


class _ContinuousSoftmaxShim:
    """Forward of deep-spin/infinite-former ``continuous_softmax.py`` (un-vendored):
    canonical parameters theta=[mu/sigma^2, -1/(2 sigma^2)] -> E_{N(mu,sigma^2)}[psi_j]."""

    def __init__(self, psi=None):
        self.psi = psi

    def __call__(self, theta):
        sigma_sq = -0.5 / theta[:, 1]
        mu = theta[:, 0] * sigma_sq
        return self.psi[0].integrate_psi_gaussian(mu.unsqueeze(1), sigma_sq.unsqueeze(1))


def load_gibbs_vl():
    """Reference live variant, Video-LLaMA flavour (T=32, e=768 hard-coded)."""
    m = _load_pkg("_ref_vl", _VL, ["basis_functions", "long_term_attention_gibbs"])
    return m["long_term_attention_gibbs"]


def load_gibbs_vc():
    """Reference live variant, VideoChat2 flavour (14x14 tokens, e=1024 hard-coded)."""
    m = _load_pkg("_ref_vc", _VC, ["basis_functions", "long_term_attention_gibbs"])
    return m["long_term_attention_gibbs"]


def load_gaussian_vl():
    """Reference Gaussian/closed-form variant (dead code upstream; needs the shim)."""
    m = _load_pkg("_ref_vl", _VL, ["basis_functions", "long_term_attention"])
    mod = m["long_term_attention"]
    mod.ContinuousSoftmax = _ContinuousSoftmaxShim
    return mod


def load_qformer_vl(ltm_module=None, pkg="_ref_vl"):
    """Reference caller: ``Qformer.py`` of Video-LLaMA (``BertSelfAttention`` :115-310 constructs, triggers and
    blends the LTM).  Written for transformers 4.x; three helpers it imports from ``transformers.modeling_utils``
    moved / disappeared in 5.x and are aliased or stubbed (they are only used by ``prune_heads``), and its absolute
    import of the LTM module (:50) is pointed at the isolated copy loaded above -- or, for the import-swap test, at
    `ltm_module` (any module exposing ``LongTermAttention``), in which case the file is loaded a second time under
    the package name `pkg` so that both flavours can live side by side."""
    import transformers.modeling_utils as MU
    import transformers.pytorch_utils as PU

    def _missing(name):
        def f(*a, **k):
            raise NotImplementedError(name)
        return f
    for n in ("apply_chunking_to_forward", "find_pruneable_heads_and_indices", "prune_linear_layer"):
        if not hasattr(MU, n):
            setattr(MU, n, getattr(PU, n, None) or _missing(n))
    ltm = load_gibbs_vl() if ltm_module is None else ltm_module
    for name in ("InfVideoLLaMA", "InfVideoLLaMA.models"):
        sys.modules.setdefault(name, types.ModuleType(name))
    alias = "InfVideoLLaMA.models.long_term_attention_gibbs"
    saved = sys.modules.get(alias)
    sys.modules[alias] = ltm
    try:
        return _load_pkg(pkg, _VL, ["Qformer"])["Qformer"]
    finally:
        if saved is not None:
            sys.modules[alias] = saved


def bert_config(num_basis, tau, alpha, sticky=True, encoder_width=768):
    """The HF config fields the video Q-former is built with (infinityqa.py:36-55)."""
    from transformers.models.bert.configuration_bert import BertConfig
    cfg = BertConfig(hidden_size=768, num_attention_heads=12)
    cfg.encoder_width = encoder_width
    cfg.alpha, cfg.num_basis, cfg.sticky, cfg.sigmas, cfg.tau = alpha, num_basis, sticky, None, tau
    cfg.attention_probs_dropout_prob = 0.0
    return cfg


def caller_kwargs(num_basis, tau, sticky, proj_key, proj_value, sigmas=None, n_heads=12, head_size=64):
    """Keyword set the reference caller passes (Qformer.py:135-158)."""
    return dict(attn_num_basis=num_basis, head_size=head_size, length=768, target_len=768,
                attn_func="softmax", infinite_memory=True, n_layers=2, attn_drop=0.1,
                n_heads=n_heads, d_model=n_heads * head_size, affines=True, mask=True,
                mask_type="cnn", kl_regularizer=False, sigma_0=None, mu_0=None,
                sticky_memories=sticky, continuous=True, sigmas=sigmas, tau=tau,
                proj_key=proj_key, proj_value=proj_value)
