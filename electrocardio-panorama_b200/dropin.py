"""Puts the B200 implementation behind the reference's ``network`` import name inside a reference process.

The reference's drivers reach the hot path through one import, ``from network import build_model, build_loss``
(codes/solver/solver.py:10); every other package they import (``solver``, ``utils``, ``dataset``, ``config``) must stay
the reference's own.  ``python main.py`` puts ``codes/`` first on ``sys.path``, so a ``PYTHONPATH`` entry cannot shadow
``codes/network`` -- and this package's ``utils`` / ``dataset`` directories must not shadow the reference's either.
``install()`` therefore registers this package's ``network`` directly in ``sys.modules`` (its own imports are
relative) and exposes the two optional device-side callers under names that collide with nothing:

    import dropin; dropin.install()          # before the reference imports `network`
    from network import build_model           # -> electrocardio-panorama_b200/network
    from utils import CheckPointer            # -> the reference's codes/utils, untouched
    import nefnet_b200_mertic                 # PSNR / PsnrAccumulator   (utils/mertic.py of this package)
    import nefnet_b200_tianchi                # pack_records / prepare_segments (dataset/tianchi.py of this package)

Launcher form -- runs an unmodified reference script with the substitution in place:

    cd /path/to/Electrocardio-Panorama/codes
    CUDA_VISIBLE_DEVICES=0 python /path/to/repo/electrocardio-panorama_b200/dropin.py main.py --config-file config/nef_net.yml

Data parallel (one process per GPU; the solver's nn.DataParallel branch, solver.py:32-34, is never taken because each rank
sees ONE device): launched under torchrun the launcher pins the rank to its GPU (CUDA_VISIBLE_DEVICES = LOCAL_RANK) and
initialises the NCCL process group before the script starts; the module then averages the gradients inside backward().

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        /path/to/repo/electrocardio-panorama_b200/dropin.py main.py --config-file config/nef_net.yml
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
_ALIASES = (("nefnet_b200_mertic", os.path.join("utils", "mertic.py")),
            ("nefnet_b200_tianchi", os.path.join("dataset", "tianchi.py")))


def install():
    """Idempotent.  Raises if a different ``network`` package was imported first (too late to substitute)."""
    net_dir = os.path.join(HERE, "network")
    have = sys.modules.get("network")
    if have is not None:
        if os.path.dirname(os.path.abspath(getattr(have, "__file__", "") or "")) != net_dir:
            raise RuntimeError("dropin.install(): another `network` package (%s) is already imported; call install() "
                               "before the reference imports it" % getattr(have, "__file__", "?"))
    else:
        spec = importlib.util.spec_from_file_location("network", os.path.join(net_dir, "__init__.py"),
                                                      submodule_search_locations=[net_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["network"] = mod
        try:
            spec.loader.exec_module(mod)
        except BaseException:
            del sys.modules["network"]
            raise
    for alias, rel in _ALIASES:
        if alias not in sys.modules:
            spec = importlib.util.spec_from_file_location(alias, os.path.join(HERE, rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[alias] = mod
            try:
                spec.loader.exec_module(mod)   # `from network import _native` inside resolves to the module above
            except BaseException:
                del sys.modules[alias]
                raise
    return sys.modules["network"]


def _init_data_parallel():
    """Under torchrun (WORLD_SIZE > 1): one visible device per rank, NCCL process group up before the reference's script runs."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return False
    local = os.environ.get("LOCAL_RANK", "0")
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    devs = visible.split(",") if visible else None
    os.environ["CUDA_VISIBLE_DEVICES"] = devs[int(local)] if devs and int(local) < len(devs) else local
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        torch.cuda.set_device(0)
        dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
    return True


def main(argv):
    if not argv:
        sys.stderr.write(__doc__)
        return 2
    import runpy
    script = os.path.abspath(argv[0])
    _init_data_parallel()
    install()
    sys.argv = [script] + list(argv[1:])
    sys.path.insert(0, os.path.dirname(script))   # what `python script.py` does
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    if sys.path and os.path.abspath(sys.path[0]) == HERE:
        sys.path.pop(0)   # this package's utils/ and dataset/ must not shadow the reference's
    sys.exit(main(sys.argv[1:]))
