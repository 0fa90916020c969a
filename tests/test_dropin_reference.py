"""CPU, build container only (needs /root/reference): the UNMODIFIED reference Solver, CheckPointer and optimiser factory
drive this package's `network` after `dropin.install()` -- SURVEY 8(b) "so codes/train_net.py and codes/val_net.py drive it
unchanged".  Without a GPU the run stops exactly where the hot path starts: inside Model_nefnet.forward, with the loud
"no CPU path" error.  Third-party modules the reference imports at module scope but this image lacks (tensorboardX,
matplotlib, skimage, yacs; SURVEY 8b "harness shims") are stubbed by the TEST, in the child process only."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "electrocardio-panorama_b200")
REF = os.environ.get("NEF_REFERENCE", "/root/reference/codes")

CHILD = textwrap.dedent(r'''
    import importlib, os, sys, tempfile, types
    PKG, REF = sys.argv[1], sys.argv[2]

    def stub(name, **attrs):
        try:
            importlib.import_module(name)
            return
        except ImportError:
            pass
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            n = ".".join(parts[:i])
            if n not in sys.modules:
                m = types.ModuleType(n); m.__path__ = []
                sys.modules[n] = m
                if i > 1:
                    setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)
        for k, v in attrs.items():
            setattr(sys.modules[name], k, v)

    stub("tensorboardX", SummaryWriter=object)
    stub("matplotlib.pyplot", cm=types.SimpleNamespace(Blues=None))
    stub("skimage.metrics", structural_similarity=lambda *a, **k: float("nan"))
    stub("setproctitle", setproctitle=lambda s: None)

    class Node(dict):          # the few yacs.config.CfgNode features config/default.py and the solver use
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    stub("yacs.config", CfgNode=Node)

    sys.path.insert(0, REF)                       # what `python main.py` from codes/ does
    sys.path.append(PKG)                          # only so that `import dropin` resolves; it is LAST
    import dropin
    net = dropin.install()
    assert dropin.install() is net                # idempotent

    import torch
    from config import cfg                        # the reference's defaults (config/default.py)
    cfg.MODEL.model = "model_nefnet"              # config/nef_net.yml
    cfg.MODEL.theta_L = 1
    cfg.DATA.lead_num = 3
    cfg.SOLVER.loss_factor = [0.5, 0.5, 1]
    cfg.SOLVER.lr = 1e-1
    cfg.SOLVER.scheduler = "MultiStep"
    cfg.SOLVER.lr_step = [50, 100]
    out = tempfile.mkdtemp()
    cfg.output_dir = out
    cfg.desc = "debug"
    os.makedirs(os.path.join(out, "debug"))

    from solver import Solver                     # the reference's, unmodified
    from solver.optim_scheduler import get_lr_scheduler, get_optimizer
    from utils import CheckPointer, seed_torch
    import solver as ref_solver, utils as ref_utils, network
    assert os.path.abspath(ref_solver.__file__).startswith(REF) and os.path.abspath(ref_utils.__file__).startswith(REF)
    assert os.path.abspath(network.__file__).startswith(PKG), network.__file__
    assert "dataset" not in sys.modules or os.path.abspath(sys.modules["dataset"].__file__).startswith(REF)

    seed_torch(seed=cfg.seed)
    s = Solver(cfg, use_tensorboardx=False)       # build_model(cfg).float(), build_loss(cfg), .to(device)
    assert type(s.model).__module__ == "network.model_nefnet" and s.loss is network.losswrapper
    assert sum(p.numel() for p in s.model.parameters()) == 7626369      # SURVEY 8(b), G = 3
    optim = get_optimizer(cfg, s.model.parameters())
    sched = get_lr_scheduler(cfg, optim)
    ck = CheckPointer(s.model, optim, sched, s.output_dir)
    assert ck.load(cfg.MODEL.resume) == {}
    ck.save("epoch_0", epoch=0, best_test_loss=1.0)
    w = s.model.state_dict()["mlp1.weight"].clone()
    with torch.no_grad():
        s.model.mlp1.weight.add_(1.0)
    extra = ck.load()                             # last_checkpoint -> epoch_0.pkl
    assert extra["epoch"] == 0 and torch.equal(s.model.state_dict()["mlp1.weight"], w)

    B, G, L, V = 2, 3, 512, 8
    meta = dict(data=torch.rand(B, G, L), rois=torch.tensor([[[0, 64], [64, 128], [128, 192], [192, 256], [256, 320],
                [320, 384], [384, 512]]] * B), input_theta=torch.rand(B, G, 2), target_view=torch.rand(B, L),
                target_theta=torch.rand(B, 2), ori_data=torch.rand(B, 12, L), noise=torch.zeros(B, L),
                rest_view=torch.rand(B, V, L), rest_theta=torch.rand(B, V, 2), unsupervision_lead_name=["v4"] * B)
    try:
        s.run_one_epoch([meta], "train", optim)   # solver.py:141-235, unmodified
    except RuntimeError as e:
        if torch.cuda.is_available():
            raise
        assert "no CPU path" in str(e) or "CUDA" in str(e), e
        print("REACHED_HOT_PATH")
    else:
        assert torch.cuda.is_available()
        print("RAN_ON_GPU")
    import nefnet_b200_mertic, nefnet_b200_tianchi
    assert hasattr(nefnet_b200_mertic, "PsnrAccumulator") and hasattr(nefnet_b200_tianchi, "prepare_segments")
    print("DROPIN_OK")
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout exists only in the build container")
def test_unmodified_reference_solver_drives_the_b200_network(tmp_path):
    script = tmp_path / "child.py"
    script.write_text(CHILD)
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, str(script), PKG, REF], capture_output=True, text=True, timeout=600, env=env,
                       cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + "\n" + r.stderr[-4000:]
    assert "DROPIN_OK" in r.stdout and ("REACHED_HOT_PATH" in r.stdout or "RAN_ON_GPU" in r.stdout)


def test_install_refuses_a_foreign_network_package(tmp_path):
    """Too late to substitute once the reference's own `network` is imported: loud, not silent."""
    (tmp_path / "network").mkdir()
    (tmp_path / "network" / "__init__.py").write_text("x = 1\n")
    code = ("import sys; sys.path.insert(0, %r); import network; sys.path.append(%r); import dropin\n"
            "try:\n    dropin.install()\nexcept RuntimeError as e:\n    print('REFUSED', e)\n" % (str(tmp_path), PKG))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "REFUSED" in r.stdout, r.stdout + r.stderr


def test_launcher_substitutes_network_only(tmp_path):
    """`python dropin.py main.py ...` from a codes/-shaped directory: `network` is this package's, `utils` / `dataset` stay the
    directory's own (this package has same-named sub-packages that must not shadow them), argv is passed through."""
    codes = tmp_path / "codes"
    for d, body in (("utils", "X = 1\n"), ("dataset", "Y = 2\n"), ("network", "raise SystemExit('the reference network was imported')\n")):
        (codes / d).mkdir(parents=True)
        (codes / d / "__init__.py").write_text(body)
    (codes / "main.py").write_text("import sys\nimport network, utils, dataset\nimport nefnet_b200_mertic, nefnet_b200_tianchi\n"
                                   "print(network.__file__, utils.X, dataset.Y, sys.argv[1:], __name__)\n")
    r = subprocess.run([sys.executable, os.path.join(PKG, "dropin.py"), "main.py", "--config-file", "config/nef_net.yml"],
                       capture_output=True, text=True, timeout=300, cwd=str(codes))
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.join(PKG, "network") in r.stdout and " 1 2 ['--config-file', 'config/nef_net.yml'] __main__" in r.stdout


def test_in_package_backend_switch(tmp_path):
    """The reference-side variant INTEGRATION.md shows: a few lines at the top of codes/network/__init__.py hand the import
    over to this package when NEFNET_BACKEND=b200 (Python re-reads sys.modules after the package body ran)."""
    codes = tmp_path / "codes"
    (codes / "network").mkdir(parents=True)
    (codes / "utils").mkdir()
    (codes / "utils" / "__init__.py").write_text("X = 1\n")
    (codes / "network" / "__init__.py").write_text(
        "import os, sys\n"
        "if os.environ.get('NEFNET_BACKEND') == 'b200':\n"
        "    sys.path.append(os.environ['NEFNET_B200_DIR'])\n"
        "    del sys.modules['network']\n"
        "    import dropin; dropin.install()\n"
        "    from network import build_model, build_loss\n"
        "else:\n"
        "    def build_model(cfg):\n        return 'reference'\n")
    (codes / "main.py").write_text("from network import build_model\nimport network, utils\n"
                                   "print(network.__file__, build_model.__module__, utils.X)\n")
    env = dict(os.environ, NEFNET_BACKEND="b200", NEFNET_B200_DIR=PKG)
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, "main.py"], capture_output=True, text=True, timeout=300, cwd=str(codes), env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.join(PKG, "network") in r.stdout and " network 1" in r.stdout
    env.pop("NEFNET_BACKEND")
    r = subprocess.run([sys.executable, "main.py"], capture_output=True, text=True, timeout=300, cwd=str(codes), env=env)
    assert r.returncode == 0 and str(codes) in r.stdout            # without the switch the tree's own package is used
