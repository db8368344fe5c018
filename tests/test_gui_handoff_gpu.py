"""f-4 (SURVEY.md section 8 f): the CUDA VolumeRenderer under the call pattern of the reference's matplotlib front
end (pyvr/interface/matplotlib_interface.py): constructed with fast() + a directional light, camera set BEFORE the
volume is loaded (:88-93), transfer functions pushed when dirty (:121-126), per frame a camera-linked light update +
set_light (:223-228), set_camera + render_to_pil (:181-185), and the switch to RenderConfig.fast() while dragging
and back afterwards (:356-401).  matplotlib itself is not needed: the loop below is what its event handlers do."""

import numpy as np
import pytest

import oracle
from pyvr_b200 import (Camera, ColorTransferFunction, Light, OpacityTransferFunction, RenderConfig, Volume,
                       build_rgba_lut, create_sample_volume)
from pyvr_b200.cuda_renderer import VolumeRenderer

from scenes import assert_parity

pytestmark = pytest.mark.gpu


def test_interactive_session_call_pattern():
    data = create_sample_volume(64, "double_sphere")
    volume = Volume(data=data, normals=oracle.normals(data))
    camera = Camera.isometric_view(distance=3.0)
    light = Light.camera_linked()
    width, height = 256, 192
    renderer = VolumeRenderer(width=width, height=height, config=RenderConfig.fast(), light=light)
    renderer.set_camera(camera)                       # before load_volume, as InteractiveVolumeRenderer.__init__ does
    renderer.load_volume(volume)
    ctf, otf = ColorTransferFunction.from_colormap("viridis"), OpacityTransferFunction.linear(0.0, 0.3)
    renderer.set_transfer_functions(ctf, otf)
    saved = RenderConfig.high_quality()
    renderer.set_config(saved)
    frames = []
    for k in range(6):                                # a drag: orbit the camera, fast preset, light follows the camera
        if k == 1:
            renderer.set_config(RenderConfig.fast())              # _switch_to_interaction_quality
        if k == 4:
            renderer.set_config(saved)                            # _restore_quality_after_interaction
            otf = OpacityTransferFunction.linear(0.0, 0.6)        # the user edited the opacity curve
            renderer.set_transfer_functions(ctf, otf)
        cam = Camera.from_spherical(target=np.zeros(3, np.float32), azimuth=0.6 + 0.25 * k, elevation=0.5,
                                    roll=0.0, distance=3.0)
        lt = renderer.get_light()
        assert lt.is_linked
        lt.update_from_camera(cam)
        renderer.set_light(lt)
        renderer.set_camera(cam)
        image = renderer.render_to_pil()
        arr = np.array(image)
        assert arr.shape == (height, width, 4) and arr.dtype == np.uint8 and arr[..., 3].max() > 50
        frames.append(arr)
        # every frame equals what the oracle renders for the state the GUI believes it has set
        cfg = renderer.get_config()
        want, _, _ = oracle.render(volume, cam, lt, cfg, build_rgba_lut(ctf, otf), width, height)
        assert_parity(arr[::-1], want)                            # render_to_pil flips to top-down
    assert renderer.get_config() is saved and renderer.get_camera() is not None and renderer.get_volume() is volume
    assert any(not np.array_equal(frames[0], f) for f in frames[1:])
    renderer.close()
