"""Locates the unmodified reference (PhiML) for tests that drive the plug-in through PhiML's own front end:
``baseline/_ref`` (pip-installed copy that ships to the GPU box, see __graft_entry__.install_reference) or, in the authoring
container, /root/reference.  Nothing here is used by the product path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference')


def reference_path():
    for p in CANDIDATES:
        if os.path.isdir(os.path.join(p, 'phiml')):
            return p
    return None


def import_phiml():
    """Returns the ``phiml`` module or None when no copy of the reference is available."""
    p = reference_path()
    if p is None:
        return None
    if p not in sys.path:
        sys.path.insert(0, p)
    import phiml
    return phiml
