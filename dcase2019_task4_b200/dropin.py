"""Run the reference's own scripts on this package: ``install()`` registers the package's modules in ``sys.modules``
under the names ``baseline/main.py:19-29`` imports (``DataLoad``, ``models.CRNN``, ``utils.utils``, ``utils.Scaler``,
``utils.ramps``, ``utils.Logger``, ``config``, ``evaluation_measures``, ``DatasetDcase2019Task4``), so that

    python -c "from dcase2019_task4_b200 import dropin; dropin.install(); import runpy; runpy.run_path('main.py', run_name='__main__')"

executed inside ``baseline/`` runs ``main.py`` / ``main_simple_CRNN.py`` / ``TestModel.py`` UNCHANGED: their own ``train``
(main.py:52-165) then drives ``models.CRNN.CRNN`` through autograd (``dcase_crnn_forward`` / ``dcase_crnn_backward``),
the transform chain runs as ``dcase_logmel_finish`` and the features come from ``dcase_logmel_fwd``; swapping in this
package's fused ``train`` is the one-line change ``from dcase2019_task4_b200.main import train``.

Not provided: the youtube download (``download_data.py``); ``initialize_and_get_df(..., download=True)`` only checks
that the audio is already on disk.
"""
import importlib
import sys

ALIASES = {
    "config": "config",
    "DataLoad": "DataLoad",
    "DatasetDcase2019Task4": "DatasetDcase2019Task4",
    "evaluation_measures": "evaluation_measures",
    "models": "models",
    "models.CNN": "models.CNN",
    "models.RNN": "models.RNN",
    "models.CRNN": "models.CRNN",
    "utils": "utils",
    "utils.utils": "utils.utils",
    "utils.Scaler": "utils.Scaler",
    "utils.ramps": "utils.ramps",
    "utils.Logger": "utils.Logger",
}


def install(force=False):
    """Alias the package's modules under the reference's top-level names.  Refuses to shadow an already imported module
    of the same name that is not ours unless ``force``."""
    pkg = __name__.rsplit(".", 1)[0]
    for alias, rel in ALIASES.items():
        mod = importlib.import_module(pkg + "." + rel)
        have = sys.modules.get(alias)
        if have is not None and have is not mod and not force:
            raise ImportError("a different module named %r is already imported (%s); call install(force=True) to "
                              "replace it" % (alias, getattr(have, "__file__", "?")))
        sys.modules[alias] = mod
    return sorted(ALIASES)


def uninstall():
    pkg = __name__.rsplit(".", 1)[0]
    for alias, rel in ALIASES.items():
        if sys.modules.get(alias) is sys.modules.get(pkg + "." + rel):
            sys.modules.pop(alias, None)
