"""OUT OF SCOPE: the reference's cv2 debug drawing.  The adapters import
``plot_box`` at module load (byte_tracker.py:21) but only call it with --online-visualization."""


def plot_box(*args, **kwargs):
    raise NotImplementedError("busca_b200 does not ship the debug GUI (busca/visualization.py is out of scope)")


def create_batch_image(*args, **kwargs):
    raise NotImplementedError("busca_b200 does not ship the debug GUI (busca/visualization.py is out of scope)")
