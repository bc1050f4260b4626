"""Launcher that runs an UNCHANGED reference script on the native path:

    python -m lavender_b200.run /path/to/LAVENDER/main_pretrain_mlm.py --config _args/args_pretrain_webvid.json ...
    python -m torch.distributed.run --nproc-per-node 8 -m lavender_b200.run /path/to/LAVENDER/main_pretrain_mlm.py ...

It puts the drop-in overlay (dropin/: utils/lib.py, utils/dist.py, utils/deepspeed.py, visbackbone/video_swin.py,
model.py, agent.py) ahead of the script's own directory on sys.path, changes into the script's directory (the
reference resolves ./_args, ./_models, ./visbackbone relative to CWD: video_swin.py:574-593, utils/args.py:3-5) and
executes the script as __main__."""
import os
import runpy
import sys


def overlay_dir():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin")


def install_overlay(script_dir=None):
    repo = os.path.dirname(overlay_dir())
    for p in (script_dir, repo, overlay_dir()):
        if p:
            while p in sys.path:
                sys.path.remove(p)
            sys.path.insert(0, p)
    for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.") or m in ("model", "agent")]:
        del sys.modules[name]


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    sdir = os.path.dirname(script)
    install_overlay(sdir)
    os.chdir(sdir)
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
