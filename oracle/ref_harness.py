"""Runs the UNMODIFIED reference modules from /root/reference on CPU -- TEST
INFRASTRUCTURE, authoring container only (the GPU box has no /root/reference; nothing
under tests/ -m gpu, smoke() or bench.py imports this file).

Three gaps are filled so that lib/models/{tepose,spin,smpl}.py and
lib/utils/geometry.py import as they are:
  1. ``yacs`` is not installed  -> a minimal ``yacs.config.CfgNode`` stub
     (lib/core/config.py:19 only builds defaults with it);
  2. ``smplx`` is not installed / not vendored (requirements.txt:7) -> a stand-in
     module exposing ``SMPL``, ``body_models.SMPLOutput`` and ``lbs.vertices2joints``
     (the three names lib/models/smpl.py:7-9 imports) implemented with the oracle's
     restated LBS.  This is why the SMPL part of the goldens is "parity unpinned";
  3. licensed assets are absent -> synthetic ``data/base_data`` in a temp cwd
     (paths are cwd-relative: lib/core/config.py:31).
"""
from __future__ import annotations

import contextlib
import os
import pickle
import sys
import tempfile
import types
from collections import namedtuple

import numpy as np
import torch

from . import synth, torch_ref

# The reference tree: /root/reference in the authoring container.  On the GPU box that path does not exist; what travels there is
# oracle/_ref/ (git-ignored, NOT gpurun-ignored): the handful of UNMODIFIED reference module files of the path, staged by
# stage_reference() from __graft_entry__.build() -- the Python counterpart of a compiled oracle/_ref/*.so.  Only bench.py's
# --impl reference arm and its cpu_baseline leg execute them there.
_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref")
STAGED_FILES = ("lib/models/tepose.py", "lib/models/spin.py", "lib/models/smpl.py", "lib/utils/geometry.py", "lib/core/config.py")


def _pick_root() -> str:
    if os.environ.get("TEPOSE_REF_ROOT"):                       # tests: force the staged copy
        return os.environ["TEPOSE_REF_ROOT"]
    if os.path.isfile(os.path.join("/root/reference", "lib", "models", "tepose.py")):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "models", "tepose.py"))


def stage_reference(src: str = "/root/reference") -> bool:
    """Copies the path's reference modules, byte for byte, into oracle/_ref/ (outputs only under oracle/_ref/, which is
    git-ignored: reference sources never enter the repository's history).  Returns False when `src` is absent."""
    import shutil
    if not os.path.isfile(os.path.join(src, "lib", "models", "tepose.py")):
        return False
    for rel in STAGED_FILES:
        dst = os.path.join(STAGED_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
    with open(os.path.join(STAGED_ROOT, "README"), "w") as fh:
        fh.write("Unmodified copies of the reference's hot-path modules (staged by oracle/ref_harness.stage_reference from "
                 "/root/reference); executed only by bench.py --impl reference / cpu_baseline.  Not tracked by git.\n")
    return True


def _install_yacs_stub():
    if "yacs.config" in sys.modules:
        return

    class CfgNode(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

        def clone(self):
            import copy
            return copy.deepcopy(self)

        def merge_from_file(self, path):
            raise NotImplementedError("yacs stub")

    yacs = types.ModuleType("yacs")
    cfgmod = types.ModuleType("yacs.config")
    cfgmod.CfgNode = CfgNode
    yacs.config = cfgmod
    sys.modules["yacs"] = yacs
    sys.modules["yacs.config"] = cfgmod


def _install_smplx_standin():
    if "smplx" in sys.modules and getattr(sys.modules["smplx"], "_tepose_b200_standin", False):
        return
    SMPLOutput = namedtuple("SMPLOutput", ["vertices", "joints", "full_pose", "betas",
                                           "global_orient", "body_pose"])
    SMPLOutput.__new__.__defaults__ = (None,) * 6

    class SMPL(torch.nn.Module):
        """smplx.SMPL stand-in: same constructor keywords the reference uses
        (lib/models/spin.py:226-230, evaluate.py:130-135), same buffers / Parameters
        (SURVEY.md App. A.1), forward restated in oracle/torch_ref.py:smpl_lbs."""

        def __init__(self, model_path, batch_size=1, create_transl=True, gender="neutral", **kw):
            super().__init__()
            fn = os.path.join(model_path, f"SMPL_{gender.upper()}.pkl") if os.path.isdir(model_path) else model_path
            with open(fn, "rb") as fh:
                data = pickle.load(fh, encoding="latin1")
            m = torch_ref.SmplModel(data, {"J_regressor_extra": np.zeros((9, 6890), np.float32),
                                           "J_regressor_h36m": np.zeros((17, 6890), np.float32)})
            self._m = m
            self.faces = m.faces
            self.register_buffer("faces_tensor", torch.as_tensor(m.faces))
            self.register_buffer("v_template", m.v_template)
            self.register_buffer("shapedirs", m.shapedirs)
            self.register_buffer("J_regressor", m.J_regressor)
            self.register_buffer("posedirs", m.posedirs)
            self.register_buffer("parents", m.parents)
            self.register_buffer("lbs_weights", m.lbs_weights)
            self.betas = torch.nn.Parameter(torch.zeros(batch_size, 10))
            self.global_orient = torch.nn.Parameter(torch.zeros(batch_size, 3))
            self.body_pose = torch.nn.Parameter(torch.zeros(batch_size, 69))

        def forward(self, betas=None, body_pose=None, global_orient=None, transl=None,
                    return_verts=True, return_full_pose=False, pose2rot=True, **kw):
            m = self._m
            if pose2rot:
                full = torch.cat([global_orient.reshape(-1, 3), body_pose.reshape(-1, 69)], dim=1)
                R = torch_ref.batch_rodrigues_smplx(full.reshape(-1, 3)).reshape(-1, 24, 3, 3)
            else:
                full = torch.cat([global_orient.reshape(-1, 1, 3, 3), body_pose.reshape(-1, 23, 3, 3)], dim=1)
                R = full
            verts, posed_J = torch_ref.smpl_lbs(m, betas, R)
            joints = torch.cat([posed_J, verts[:, m.extra_vertex_ids]], dim=1)
            return SMPLOutput(vertices=verts, joints=joints, full_pose=full, betas=betas,
                              global_orient=global_orient, body_pose=body_pose)

    def vertices2joints(J_regressor, vertices):
        return torch.einsum("bik,ji->bjk", vertices, J_regressor)

    smplx = types.ModuleType("smplx")
    smplx._tepose_b200_standin = True
    smplx.SMPL = SMPL
    bm = types.ModuleType("smplx.body_models")
    bm.SMPLOutput = SMPLOutput
    lbs = types.ModuleType("smplx.lbs")
    lbs.vertices2joints = vertices2joints
    smplx.body_models, smplx.lbs = bm, lbs
    sys.modules.update({"smplx": smplx, "smplx.body_models": bm, "smplx.lbs": lbs})


def _install_lib_packages():
    """Register ``lib``, ``lib.models`` ... as bare namespace packages so that
    lib/models/__init__.py (which drags in the discriminator) is not executed; the
    hot-path module files themselves are imported unmodified."""
    for name in ("lib", "lib.core", "lib.models", "lib.utils"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = [os.path.join(REFERENCE_ROOT, *name.split("."))]
            sys.modules[name] = mod


@contextlib.contextmanager
def reference_env(seed: int = 0):
    """cwd = temp dir holding synthetic data/base_data; yields the reference modules."""
    if not available():
        raise RuntimeError("/root/reference is not present on this box")
    _install_yacs_stub()
    _install_smplx_standin()
    _install_lib_packages()
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_base_data(os.path.join(tmp, "data", "base_data"), seed)
        os.chdir(tmp)
        try:
            import importlib
            mods = types.SimpleNamespace(
                tepose=importlib.import_module("lib.models.tepose"),
                spin=importlib.import_module("lib.models.spin"),
                smpl=importlib.import_module("lib.models.smpl"),
                geometry=importlib.import_module("lib.utils.geometry"),
            )
            yield mods
        finally:
            os.chdir(old)


def build_reference_model(mods, sd_np: dict, seqlen: int, n_layers: int, hidden: int):
    """lib.models.TePose with our synthetic parameters loaded (strict on everything
    that make_state_dict provides; regressor.smpl.* keeps what the constructor read)."""
    model = mods.tepose.TePose(seqlen=seqlen, n_layers=n_layers, hidden_size=hidden, pretrained="")
    own = model.state_dict()
    for k, v in sd_np.items():
        assert k in own and tuple(own[k].shape) == tuple(v.shape), k
        own[k] = torch.as_tensor(v)
    missing = [k for k in own if k not in sd_np and not k.startswith("regressor.smpl.")]
    assert not missing, missing
    model.load_state_dict(own, strict=True)
    return model.eval()


def run_reference(seed, batch, seqlen, n_layers, hidden, is_train=False, use_h36m=False, x=None):
    sd = synth.make_state_dict(seed, n_layers, hidden)
    if x is None:
        x = synth.make_input(seed, batch, seqlen)
    with reference_env(seed) as mods:
        model = build_reference_model(mods, sd, seqlen, n_layers, hidden)
        Jr = None
        if use_h36m:
            Jr = torch.from_numpy(np.load(os.path.join("data", "base_data", "J_regressor_h36m.npy"))).float()
        with torch.no_grad():
            out = model(torch.from_numpy(x), is_train=is_train, J_regressor=Jr)[-1]
    return {k: v.detach().numpy().copy() for k, v in out.items()}


def run_reference_train(seed, batch, seqlen, n_layers, hidden):
    """The unmodified reference in TRAIN mode (lib/core/trainer.py:137,203: `generator.train()`,
    `generator(inp, is_train=True)`), one synthetic-loss backward.  nn.Dropout draws its masks from torch's RNG; to make the
    run reproducible from numpy seeds the given masks are forced through forward hooks on `regressor.drop1/drop2`
    (the hook returns input * mask / (1 - p), i.e. what nn.Dropout computes for that mask).  Returns outputs, loss and
    every parameter's gradient."""
    from . import train_ref
    sd = synth.make_state_dict(seed, n_layers, hidden)
    x = synth.make_input(seed, batch, seqlen)
    masks = torch.from_numpy(train_ref.make_masks(seed, 2 * batch))
    tgt = train_ref.make_targets(seed, 2 * batch)
    with reference_env(seed) as mods:
        model = build_reference_model(mods, sd, seqlen, n_layers, hidden).train()
        calls = {"drop1": 0, "drop2": 0}

        def hook(which, col):
            def fn(module, inp, out):
                i = calls[which]
                calls[which] += 1
                return inp[0] * masks[i, col] / (1.0 - module.p)
            return fn
        model.regressor.drop1.register_forward_hook(hook("drop1", 0))
        model.regressor.drop2.register_forward_hook(hook("drop2", 1))
        out = model(torch.from_numpy(x), is_train=True)[-1]
        loss = train_ref.synthetic_loss(out, tgt)
        model.zero_grad()
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return {k: v.detach().numpy().copy() for k, v in out.items()}, float(loss.detach()), grads


def run_reference_vibe(seed, batch, seqlen, n_layers, hidden, add_linear=False, bidirectional=False, use_residual=True,
                       use_h36m=False, x=None):
    """The unmodified lib/models/vibe.py VIBE on the synthetic parameters of `seed`."""
    import importlib
    sd = synth.make_vibe_state_dict(seed, n_layers, hidden, add_linear, bidirectional)
    if x is None:
        x = synth.make_vibe_input(seed, batch, seqlen)
    with reference_env(seed):
        vibe = importlib.import_module("lib.models.vibe")
        model = vibe.VIBE(seqlen=seqlen, n_layers=n_layers, hidden_size=hidden, add_linear=add_linear,
                          bidirectional=bidirectional, use_residual=use_residual, pretrained="")
        own = model.state_dict()
        for k, v in sd.items():
            assert k in own and tuple(own[k].shape) == tuple(v.shape), k
            own[k] = torch.as_tensor(v)
        missing = [k for k in own if k not in sd and not k.startswith("regressor.smpl.")]
        assert not missing, missing
        model.load_state_dict(own, strict=True)
        model.eval()
        Jr = None
        if use_h36m:
            Jr = torch.from_numpy(np.load(os.path.join("data", "base_data", "J_regressor_h36m.npy"))).float()
        with torch.no_grad():
            out = model(torch.from_numpy(x), J_regressor=Jr)[-1]
    return {k: v.detach().numpy().copy() for k, v in out.items()}


def run_reference_stream(seed, batch, frames, seqlen, n_layers, hidden, vibe_layers, vibe_hidden, use_h36m=True):
    """demo.py:229-252 / evaluate.py:233-269 restated around the UNMODIFIED lib.models.VIBE and lib.models.TePose:
    VIBE on the first T frames, then one TePose window per frame with the predicted thetas fed back."""
    import importlib
    sd_v = synth.make_vibe_state_dict(seed, vibe_layers, vibe_hidden, True, False)
    sd = synth.make_state_dict(seed, n_layers, hidden)
    feats = torch.from_numpy(synth.make_vibe_input(seed, batch, frames))
    T = seqlen

    def load(model, sd_np):
        own = model.state_dict()
        for k, v in sd_np.items():
            assert k in own and tuple(own[k].shape) == tuple(v.shape), k
            own[k] = torch.as_tensor(v)
        model.load_state_dict(own, strict=True)
        return model.eval()

    with reference_env(seed) as mods:
        vibe = importlib.import_module("lib.models.vibe")
        model_vibe = load(vibe.VIBE(seqlen=T, n_layers=vibe_layers, hidden_size=vibe_hidden, add_linear=True,
                                    bidirectional=False, use_residual=True, pretrained=""), sd_v)
        model = build_reference_model(mods, sd, T, n_layers, hidden)
        Jr = torch.from_numpy(np.load(os.path.join("data", "base_data", "J_regressor_h36m.npy"))).float() if use_h36m else None
        keys = ("theta", "verts", "kp_2d", "kp_3d", "rotmat")
        with torch.no_grad():
            output = model_vibe(feats[:, :T], J_regressor=Jr)[-1]
            per = {k: [output[k][:, t] for t in range(T - 1)] for k in keys}
            theta_input = output["theta"][:, :T - 1].detach().clone()
            for k in range(frames - T + 1):
                inp = torch.zeros((batch, T, 2048 + 85)).float()
                inp[:, :, :2048] = feats[:, k:k + T].clone()
                inp[:, :T - 1, 2048:] = theta_input.clone()
                preds = model(inp, J_regressor=Jr, is_train=False)[-1]
                for kk in keys:
                    per[kk].append(preds[kk])
                theta_input[:, :T - 2] = theta_input[:, 1:T - 1].clone()
                theta_input[:, T - 2] = preds["theta"].clone().detach()
    return {k: torch.stack(v, dim=1).numpy().copy() for k, v in per.items()}


def reference_eval_utils():
    """lib.utils.eval_utils imported unmodified (matplotlib, used only by plot_accel, is stubbed when absent)."""
    import importlib
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot
    _install_lib_packages()
    return importlib.import_module("lib.utils.eval_utils")
