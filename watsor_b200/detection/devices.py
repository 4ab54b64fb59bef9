"""Device discovery for the B200 back-end: the counterpart of `cuda_gpus()` in
watsor/detection/devices.py:28-77, with the same precedence rules --
`CUDA_DEVICE` > `~/.cuda_device` > every visible device, and `CUDA_VISIBLE_DEVICES=""` yields
nothing (the CUDA runtime inside libwatsor_b200 honours it).  pycuda is not needed: the count
comes from `wb_device_count()`."""
import os


def b200_gpus():
    """Yields (device_index, B200ObjectDetector) for every usable device."""
    try:
        from .. import _lib
        from .b200 import B200ObjectDetector
        ndevices = _lib.device_count()
    except Exception:
        return
    if ndevices == 0:
        return
    device = os.environ.get('CUDA_DEVICE')
    if device is None:
        try:
            homedir = os.environ.get('HOME')
            assert homedir is not None
            device = open(os.path.join(homedir, '.cuda_device')).read().strip()
        except Exception:
            pass
    if device is not None:
        try:
            device = int(device)
        except Exception as e:
            raise TypeError('CUDA device number (CUDA_DEVICE or ~/.cuda_device) must be an integer') from e
        yield device, B200ObjectDetector
    else:
        for device in range(ndevices):
            yield device, B200ObjectDetector
