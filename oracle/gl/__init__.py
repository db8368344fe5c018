"""The reference's own GLSL on Mesa llvmpipe, headless.  TEST INFRASTRUCTURE ONLY (part of ``oracle/``).

``GLReference`` replays ``pyvr/moderngl_renderer/manager.py`` + ``renderer.py`` call for call -- same
textures, formats, sampler state, uniforms, blend state, draw and read-back -- through plain ctypes
OpenGL 3.3 instead of ``moderngl`` (absent from this image), and compiles the reference's two shader
files VERBATIM (``tests/golden/shaders/``, sha256 pinned in ``tests/golden/meta.json``; read from
``/root/reference/pyvr/shaders`` instead when that tree is present and identical).

The OpenGL implementation is the Mesa 18.1.9 software ``libGL`` (gallium **llvmpipe**) that ships inside the
Nsight Compute tree of this image; it emulates GLX on Xlib, and ``fakex11.c`` provides the 27 Xlib entry
points it imports so that no X server is needed.  That is the CPU baseline BASELINE.json names ("the
reference run headless ... on Mesa llvmpipe on the box's own host cores") and the real run of the shader
that pins ``oracle/pyvr_oracle.c`` (``tests/test_gl_reference.py``).

Importers allowed: ``tests/``, ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  Nothing under
``pyvr_b200/`` may import this package.
"""

from __future__ import annotations

import ctypes
import glob
import hashlib
import os
import subprocess
from ctypes import POINTER, byref, c_char_p, c_float, c_int, c_uint, c_ulong, c_void_p
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
_REF_DIR = os.path.join(os.path.dirname(_HERE), "_ref")
_SHADER_DIRS = ["/root/reference/pyvr/shaders", os.path.join(_REPO, "tests", "golden", "shaders")]
SHADER_SHA256 = {
    "volume.frag.glsl": "870a4944cee589585a5bf786bcfc0f0618e681a5a06bbfe29c5e345251d85c96",
    "volume.vert.glsl": "72b11bd2a823748fb556e653a266e9c9ca87f3df1d16fb63a202fffa93383216",
}
_MESA_GLOBS = ["/opt/nvidia/nsight-compute/*/host/linux-desktop-glibc_2_11_3-x64/Mesa/libGL.so.1",
               "/opt/nvidia/nsight-compute/*/host/*/Mesa/libGL.so.1"]

# GL / GLX enums used below
GL_TEXTURE_2D, GL_TEXTURE_3D = 0x0DE1, 0x806F
GL_RGBA8, GL_RGBA, GL_RGB, GL_RED = 0x8058, 0x1908, 0x1907, 0x1903
GL_R32F, GL_RGB32F, GL_RGBA32F = 0x822E, 0x8815, 0x8814
GL_DEPTH_COMPONENT24, GL_DEPTH_COMPONENT = 0x81A6, 0x1902
GL_UNSIGNED_BYTE, GL_UNSIGNED_INT, GL_FLOAT = 0x1401, 0x1405, 0x1406
GL_TEXTURE_MIN_FILTER, GL_TEXTURE_MAG_FILTER, GL_LINEAR = 0x2801, 0x2800, 0x2601
GL_TEXTURE_WRAP_S, GL_TEXTURE_WRAP_T, GL_TEXTURE_WRAP_R, GL_CLAMP_TO_EDGE = 0x2802, 0x2803, 0x8072, 0x812F
GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0, GL_DEPTH_ATTACHMENT, GL_FRAMEBUFFER_COMPLETE = 0x8D40, 0x8CE0, 0x8D00, 0x8CD5
GL_VERTEX_SHADER, GL_FRAGMENT_SHADER, GL_COMPILE_STATUS, GL_LINK_STATUS = 0x8B31, 0x8B30, 0x8B81, 0x8B82
GL_ARRAY_BUFFER, GL_ELEMENT_ARRAY_BUFFER, GL_STATIC_DRAW = 0x8892, 0x8893, 0x88E4
GL_TRIANGLES, GL_BLEND, GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA = 0x0004, 0x0BE2, 0x0302, 0x0303
GL_COLOR_BUFFER_BIT, GL_DEPTH_BUFFER_BIT, GL_TEXTURE0 = 0x4000, 0x0100, 0x84C0
GL_UNPACK_ALIGNMENT, GL_PACK_ALIGNMENT = 0x0CF5, 0x0D05
GL_VENDOR, GL_RENDERER, GL_VERSION = 0x1F00, 0x1F01, 0x1F02


class GLUnavailable(RuntimeError):
    """No usable software OpenGL on this machine (the Mesa library or a C compiler is missing)."""


def mesa_library() -> Optional[str]:
    for pattern in _MESA_GLOBS:
        hits = sorted(glob.glob(pattern))
        if hits:
            return hits[-1]
    return None


def build(force: bool = False) -> str:
    """Compile ``fakex11.c`` into ``oracle/_ref/libX11.so.6`` and ``libXext.so.6``; returns the directory."""
    src = os.path.join(_HERE, "fakex11.c")
    os.makedirs(_REF_DIR, exist_ok=True)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    for soname in ("libX11.so.6", "libXext.so.6"):
        out = os.path.join(_REF_DIR, soname)
        if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            subprocess.run([cc, "-O2", "-shared", "-fPIC", f"-Wl,-soname,{soname}", "-o", out, src],
                           check=True, capture_output=True, text=True)
    return _REF_DIR


def shader_sources():
    """The reference's two shader files, verbatim (sha256-checked)."""
    for d in _SHADER_DIRS:
        try:
            texts = {n: open(os.path.join(d, n), "rb").read() for n in SHADER_SHA256}
        except OSError:
            continue
        if all(hashlib.sha256(t).hexdigest() == SHADER_SHA256[n] for n, t in texts.items()):
            return texts["volume.vert.glsl"], texts["volume.frag.glsl"], d
    raise GLUnavailable("reference shader files not found or sha256 mismatch")


_libs = None


def _load():
    """dlopen the stand-in Xlib first (matched by SONAME), then Mesa's libGL."""
    global _libs
    if _libs is None:
        mesa = mesa_library()
        if mesa is None:
            raise GLUnavailable("Mesa software libGL (Nsight Compute tree) not found")
        try:
            ref = build()
        except (OSError, subprocess.CalledProcessError) as e:
            raise GLUnavailable(f"cannot build the Xlib stand-in: {e}")
        os.environ.setdefault("XLIB_NO_SHM", "1")     # display targets from malloc, not SysV shm
        x11 = ctypes.CDLL(os.path.join(ref, "libX11.so.6"), mode=ctypes.RTLD_GLOBAL)
        ctypes.CDLL(os.path.join(ref, "libXext.so.6"), mode=ctypes.RTLD_GLOBAL)
        try:
            gl = ctypes.CDLL(mesa, mode=ctypes.RTLD_GLOBAL)
        except OSError as e:
            raise GLUnavailable(f"cannot load {mesa}: {e}")
        _libs = (x11, gl, mesa)
    return _libs


class _GL:
    """Function table resolved through glXGetProcAddress (what moderngl's loader does)."""

    def __init__(self, gl):
        gl.glXGetProcAddress.restype = c_void_p
        gl.glXGetProcAddress.argtypes = [c_char_p]
        self._lib = gl
        self._cache = {}

    def fn(self, name, restype, *argtypes):
        key = name
        if key not in self._cache:
            addr = self._lib.glXGetProcAddress(name.encode())
            if not addr:
                raise GLUnavailable(f"{name} not exported by the GL library")
            self._cache[key] = ctypes.CFUNCTYPE(restype, *argtypes)(addr)
        return self._cache[key]


class GLReference:
    """``ModernGLManager`` + the GL half of ``ModernGLVolumeRenderer``, through ctypes.

    Method names and the order of GL calls follow manager.py / renderer.py (cited per method).
    One instance per process thread (GLX contexts are thread-affine, like the reference's).
    """

    def __init__(self, width=512, height=512, threads: Optional[int] = None):
        if threads is not None:
            os.environ["LP_NUM_THREADS"] = str(int(threads))   # read by llvmpipe at screen creation
        x11, gl, mesa = _load()
        self.width, self.height, self.mesa_path = int(width), int(height), mesa
        self._g = g = _GL(gl)
        # manager.py:26  moderngl.create_context(standalone=True)  -> GLX pbuffer + 3.3 core context
        x11.XOpenDisplay.restype = c_void_p
        self._dpy = c_void_p(x11.XOpenDisplay(None))
        gl.glXChooseFBConfig.restype = POINTER(c_void_p)
        gl.glXChooseFBConfig.argtypes = [c_void_p, c_int, POINTER(c_int), POINTER(c_int)]
        attribs = (c_int * 17)(0x8011, 0x1, 0x8010, 0x4, 8, 8, 9, 8, 10, 8, 11, 8, 12, 24, 5, 0, 0)
        n = c_int(0)
        configs = gl.glXChooseFBConfig(self._dpy, 0, attribs, byref(n))
        if not configs or n.value < 1:
            raise GLUnavailable("glXChooseFBConfig found no config")
        config = c_void_p(configs[0])
        gl.glXCreatePbuffer.restype = c_ulong
        gl.glXCreatePbuffer.argtypes = [c_void_p, c_void_p, POINTER(c_int)]
        self._pbuffer = gl.glXCreatePbuffer(self._dpy, config, (c_int * 5)(0x8041, 16, 0x8040, 16, 0))
        create = g.fn("glXCreateContextAttribsARB", c_void_p, c_void_p, c_void_p, c_void_p, c_int, POINTER(c_int))
        self._ctx = create(self._dpy, config, None, 1, (c_int * 7)(0x2091, 3, 0x2092, 3, 0x9126, 0x1, 0))
        if not self._ctx:
            raise GLUnavailable("OpenGL 3.3 core context creation failed")
        gl.glXMakeContextCurrent.argtypes = [c_void_p, c_ulong, c_ulong, c_void_p]
        if not gl.glXMakeContextCurrent(self._dpy, self._pbuffer, self._pbuffer, c_void_p(self._ctx)):
            raise GLUnavailable("glXMakeContextCurrent failed")
        self._glx = gl
        # moderngl creates and updates textures on a scratch unit (its `default_texture_unit`, the last one) so
        # that user bindings made with texture.use(unit) are never disturbed
        units = c_int(0)
        g.fn("glGetIntegerv", None, c_uint, POINTER(c_int))(0x8872, byref(units))   # GL_MAX_TEXTURE_IMAGE_UNITS
        self._scratch_unit = max(units.value, 16) - 1
        get_string = g.fn("glGetString", c_char_p, c_uint)
        self.info = {"vendor": get_string(GL_VENDOR).decode(), "renderer": get_string(GL_RENDERER).decode(),
                     "version": get_string(GL_VERSION).decode(), "library": mesa,
                     "threads": os.environ.get("LP_NUM_THREADS", "default (cores, max 16)")}

        # manager.py:29-31  RGBA8 colour texture + depth texture + framebuffer
        self.color_texture = self._texture2d(GL_RGBA8, GL_RGBA, GL_UNSIGNED_BYTE, None, self.width, self.height)
        self.depth_texture = self._texture2d(GL_DEPTH_COMPONENT24, GL_DEPTH_COMPONENT, GL_FLOAT, None, self.width, self.height)
        fbo = c_uint(0)
        g.fn("glGenFramebuffers", None, c_int, POINTER(c_uint))(1, byref(fbo))
        self.fbo = fbo.value
        bind_fb = g.fn("glBindFramebuffer", None, c_uint, c_uint)
        bind_fb(GL_FRAMEBUFFER, self.fbo)
        attach = g.fn("glFramebufferTexture2D", None, c_uint, c_uint, c_uint, c_uint, c_int)
        attach(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0, GL_TEXTURE_2D, self.color_texture, 0)
        attach(GL_FRAMEBUFFER, GL_DEPTH_ATTACHMENT, GL_TEXTURE_2D, self.depth_texture, 0)
        status = g.fn("glCheckFramebufferStatus", c_uint, c_uint)(GL_FRAMEBUFFER)
        if status != GL_FRAMEBUFFER_COMPLETE:
            raise GLUnavailable(f"framebuffer incomplete (0x{status:x})")
        self.program = None
        self.vao = None
        self._next_texture_unit = 0           # manager.py:40
        self._textures = []

    # ---- helpers -------------------------------------------------------------------------------
    def _texture2d(self, internal, fmt, typ, data, w, h):
        g = self._g
        tex = c_uint(0)
        g.fn("glGenTextures", None, c_int, POINTER(c_uint))(1, byref(tex))
        g.fn("glActiveTexture", None, c_uint)(GL_TEXTURE0 + self._scratch_unit)
        g.fn("glBindTexture", None, c_uint, c_uint)(GL_TEXTURE_2D, tex.value)
        g.fn("glPixelStorei", None, c_uint, c_int)(GL_UNPACK_ALIGNMENT, 1)
        g.fn("glTexImage2D", None, c_uint, c_int, c_int, c_int, c_int, c_int, c_uint, c_uint, c_void_p)(
            GL_TEXTURE_2D, 0, internal, w, h, 0, fmt, typ, data)
        par = g.fn("glTexParameteri", None, c_uint, c_uint, c_int)
        par(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR)     # moderngl's default for a level-0-only texture
        par(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR)
        return tex.value

    def _texture3d(self, shape, components, data):
        """``ctx.texture3d(shape, components, bytes, dtype='f4')`` + LINEAR + repeat_{x,y,z}=False
        (manager.py:95-101 / :123-129): (width, height, depth) = shape over the C-order bytes."""
        g = self._g
        internal, fmt = {1: (GL_R32F, GL_RED), 3: (GL_RGB32F, GL_RGB)}[components]
        tex = c_uint(0)
        g.fn("glGenTextures", None, c_int, POINTER(c_uint))(1, byref(tex))
        g.fn("glActiveTexture", None, c_uint)(GL_TEXTURE0 + self._scratch_unit)
        g.fn("glBindTexture", None, c_uint, c_uint)(GL_TEXTURE_3D, tex.value)
        g.fn("glPixelStorei", None, c_uint, c_int)(GL_UNPACK_ALIGNMENT, 1)
        g.fn("glTexImage3D", None, c_uint, c_int, c_int, c_int, c_int, c_int, c_int, c_uint, c_uint, c_void_p)(
            GL_TEXTURE_3D, 0, internal, int(shape[0]), int(shape[1]), int(shape[2]), 0, fmt, GL_FLOAT,
            data.ctypes.data_as(c_void_p))
        par = g.fn("glTexParameteri", None, c_uint, c_uint, c_int)
        par(GL_TEXTURE_3D, GL_TEXTURE_MIN_FILTER, GL_LINEAR)
        par(GL_TEXTURE_3D, GL_TEXTURE_MAG_FILTER, GL_LINEAR)
        par(GL_TEXTURE_3D, GL_TEXTURE_WRAP_S, GL_CLAMP_TO_EDGE)
        par(GL_TEXTURE_3D, GL_TEXTURE_WRAP_T, GL_CLAMP_TO_EDGE)
        par(GL_TEXTURE_3D, GL_TEXTURE_WRAP_R, GL_CLAMP_TO_EDGE)
        err = g.fn("glGetError", c_uint)()
        if err:   # Mesa 18 llvmpipe caps one texture at 1 GiB and stores RGB32F as RGBA32F: normals stop at 406^3
            nbytes = int(shape[0]) * int(shape[1]) * int(shape[2]) * (16 if components == 3 else 4)
            raise GLUnavailable(f"glTexImage3D {tuple(int(v) for v in shape)} x {components} float failed with 0x{err:x}"
                                f"{' (GL_OUT_OF_MEMORY)' if err == 0x505 else ''}: {nbytes / 2 ** 30:.2f} GiB exceeds this "
                                "llvmpipe's 1 GiB-per-texture limit")
        self._textures.append(tex.value)
        return tex.value

    def _use(self, target, tex, unit):
        """``texture.use(unit)``."""
        self._g.fn("glActiveTexture", None, c_uint)(GL_TEXTURE0 + unit)
        self._g.fn("glBindTexture", None, c_uint, c_uint)(target, tex)

    def _get_next_texture_unit(self):   # manager.py:232-236 (never reuses a unit)
        unit = self._next_texture_unit
        self._next_texture_unit += 1
        return unit

    def _location(self, name):
        return self._g.fn("glGetUniformLocation", c_int, c_uint, c_char_p)(self.program, name.encode())

    # ---- manager.py:43-75 ------------------------------------------------------------------------
    def load_shaders(self):
        g = self._g
        vert, frag, self.shader_dir = shader_sources()

        def compile_one(kind, text):
            sh = g.fn("glCreateShader", c_uint, c_uint)(kind)
            src = c_char_p(text)
            length = c_int(len(text))
            g.fn("glShaderSource", None, c_uint, c_int, POINTER(c_char_p), POINTER(c_int))(sh, 1, byref(src), byref(length))
            g.fn("glCompileShader", None, c_uint)(sh)
            ok = c_int(0)
            g.fn("glGetShaderiv", None, c_uint, c_uint, POINTER(c_int))(sh, GL_COMPILE_STATUS, byref(ok))
            if not ok.value:
                log = ctypes.create_string_buffer(8192)
                g.fn("glGetShaderInfoLog", None, c_uint, c_int, POINTER(c_int), c_char_p)(sh, 8192, None, log)
                raise RuntimeError("shader compile failed: " + log.value.decode(errors="replace"))
            return sh

        vs, fs = compile_one(GL_VERTEX_SHADER, vert), compile_one(GL_FRAGMENT_SHADER, frag)
        prog = g.fn("glCreateProgram", c_uint)()
        g.fn("glAttachShader", None, c_uint, c_uint)(prog, vs)
        g.fn("glAttachShader", None, c_uint, c_uint)(prog, fs)
        g.fn("glLinkProgram", None, c_uint)(prog)
        ok = c_int(0)
        g.fn("glGetProgramiv", None, c_uint, c_uint, POINTER(c_int))(prog, GL_LINK_STATUS, byref(ok))
        if not ok.value:
            log = ctypes.create_string_buffer(8192)
            g.fn("glGetProgramInfoLog", None, c_uint, c_int, POINTER(c_int), c_char_p)(prog, 8192, None, log)
            raise RuntimeError("program link failed: " + log.value.decode(errors="replace"))
        self.program = prog
        g.fn("glUseProgram", None, c_uint)(prog)
        # _create_fullscreen_quad, manager.py:63-75
        vertices = np.array([-1.0, -1.0, 1.0, -1.0, 1.0, 1.0, -1.0, 1.0], dtype=np.float32)
        indices = np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32)
        vao, bufs = c_uint(0), (c_uint * 2)()
        g.fn("glGenVertexArrays", None, c_int, POINTER(c_uint))(1, byref(vao))
        g.fn("glBindVertexArray", None, c_uint)(vao.value)
        g.fn("glGenBuffers", None, c_int, POINTER(c_uint))(2, bufs)
        bind = g.fn("glBindBuffer", None, c_uint, c_uint)
        data = g.fn("glBufferData", None, c_uint, ctypes.c_ssize_t, c_void_p, c_uint)
        bind(GL_ARRAY_BUFFER, bufs[0])
        data(GL_ARRAY_BUFFER, vertices.nbytes, vertices.ctypes.data_as(c_void_p), GL_STATIC_DRAW)
        bind(GL_ELEMENT_ARRAY_BUFFER, bufs[1])
        data(GL_ELEMENT_ARRAY_BUFFER, indices.nbytes, indices.ctypes.data_as(c_void_p), GL_STATIC_DRAW)
        loc = g.fn("glGetAttribLocation", c_int, c_uint, c_char_p)(prog, b"position")
        g.fn("glEnableVertexAttribArray", None, c_uint)(loc)
        g.fn("glVertexAttribPointer", None, c_uint, c_int, c_uint, ctypes.c_ubyte, c_int, c_void_p)(loc, 2, GL_FLOAT, 0, 8, None)
        self.vao = vao.value

    # ---- uniform setters, manager.py:188-210 -------------------------------------------------------
    def set_uniform_matrix(self, name, matrix):
        m = np.ascontiguousarray(matrix, dtype=np.float32)          # .write(matrix.tobytes()): 16 floats as they lie
        self._g.fn("glUniformMatrix4fv", None, c_int, c_int, ctypes.c_ubyte, c_void_p)(
            self._location(name), 1, 0, m.ctypes.data_as(c_void_p))

    def set_uniform_vector(self, name, vector):
        v = [float(x) for x in vector]
        self._g.fn("glUniform3f", None, c_int, c_float, c_float, c_float)(self._location(name), *v)

    def set_uniform_float(self, name, value):
        self._g.fn("glUniform1f", None, c_int, c_float)(self._location(name), float(value))

    def set_uniform_int(self, name, value):
        self._g.fn("glUniform1i", None, c_int, c_int)(self._location(name), int(value))

    # ---- renderer.py:107-146 -----------------------------------------------------------------------
    def load_volume(self, data, normals, min_bounds, max_bounds):
        data = np.asarray(data)
        if data.ndim != 3:
            raise ValueError("Volume data must be 3D")
        data = np.ascontiguousarray(data, dtype=np.float32)
        unit = self._get_next_texture_unit()
        self._use(GL_TEXTURE_3D, self._texture3d(data.shape, 1, data), unit)
        self.set_uniform_int("volume_texture", unit)
        self.set_uniform_vector("volume_min_bounds", tuple(min_bounds))
        self.set_uniform_vector("volume_max_bounds", tuple(max_bounds))
        if normals is not None:
            normals = np.ascontiguousarray(normals, dtype=np.float32)
            if normals.shape[-1] != 3:
                raise ValueError("Normal volume must have 3 channels (last dimension).")
            unit = self._get_next_texture_unit()
            self._use(GL_TEXTURE_3D, self._texture3d(normals.shape[:3], 3, normals), unit)
            self.set_uniform_int("normal_volume", unit)
        # no normals: `normal_volume` is left at its default, texture unit 0 (renderer.py:143-146)

    # ---- manager.py:137-186 + renderer.py:204-207 ----------------------------------------------------
    def set_lut(self, rgba_lut):
        lut = np.ascontiguousarray(rgba_lut, dtype=np.float32)
        size = lut.shape[0]
        tex = self._texture2d(GL_RGBA32F, GL_RGBA, GL_FLOAT, lut.ctypes.data_as(c_void_p), size, 1)
        par = self._g.fn("glTexParameteri", None, c_uint, c_uint, c_int)
        par(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, GL_CLAMP_TO_EDGE)
        par(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, GL_CLAMP_TO_EDGE)
        self._textures.append(tex)
        unit = self._get_next_texture_unit()
        self._use(GL_TEXTURE_2D, tex, unit)
        self.set_uniform_int("transfer_function_lut", unit)

    # ---- renderer.py:164-172, 303-316 -----------------------------------------------------------------
    def set_camera(self, view_matrix, projection_matrix, position):
        self.set_uniform_matrix("view_matrix", view_matrix)
        self.set_uniform_matrix("projection_matrix", projection_matrix)
        self.set_uniform_vector("camera_pos", tuple(position))

    def set_config(self, step_size, max_steps, reference_step_size):
        self.set_uniform_float("step_size", step_size)
        self.set_uniform_int("max_steps", max_steps)
        self.set_uniform_float("reference_step_size", reference_step_size)

    def set_light(self, ambient, diffuse, position, target):
        self.set_uniform_float("ambient_light", ambient)
        self.set_uniform_float("diffuse_light", diffuse)
        self.set_uniform_vector("light_position", tuple(position))
        self.set_uniform_vector("light_target", tuple(target))

    # ---- renderer.py:209-219, manager.py:212-230 --------------------------------------------------------
    def render(self) -> bytes:
        g = self._g
        g.fn("glBindFramebuffer", None, c_uint, c_uint)(GL_FRAMEBUFFER, self.fbo)          # fbo.use()
        g.fn("glViewport", None, c_int, c_int, c_int, c_int)(0, 0, self.width, self.height)
        g.fn("glClearColor", None, c_float, c_float, c_float, c_float)(0.0, 0.0, 0.0, 0.0)  # ctx.clear(r,g,b,a)
        g.fn("glClearDepth", None, ctypes.c_double)(1.0)
        g.fn("glClear", None, c_uint)(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT)
        g.fn("glEnable", None, c_uint)(GL_BLEND)                                           # setup_blending
        g.fn("glBlendFunc", None, c_uint, c_uint)(GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA)
        g.fn("glUseProgram", None, c_uint)(self.program)                                   # vao.render()
        g.fn("glBindVertexArray", None, c_uint)(self.vao)
        g.fn("glDrawElements", None, c_uint, c_int, c_uint, c_void_p)(GL_TRIANGLES, 6, GL_UNSIGNED_INT, None)
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)                       # fbo.read(components=4)
        g.fn("glPixelStorei", None, c_uint, c_int)(GL_PACK_ALIGNMENT, 1)
        g.fn("glReadBuffer", None, c_uint)(GL_COLOR_ATTACHMENT0)
        g.fn("glReadPixels", None, c_int, c_int, c_int, c_int, c_uint, c_uint, c_void_p)(
            0, 0, self.width, self.height, GL_RGBA, GL_UNSIGNED_BYTE, out.ctypes.data_as(c_void_p))
        err = g.fn("glGetError", c_uint)()
        if err:
            raise RuntimeError(f"OpenGL error 0x{err:x} during render")
        return out.tobytes()

    def close(self):
        """Release the context (manager.py:238-256).  The Mesa library itself stays loaded."""
        if getattr(self, "_ctx", None):
            self._glx.glXMakeContextCurrent(self._dpy, 0, 0, None)
            self._glx.glXDestroyContext.argtypes = [c_void_p, c_void_p]
            self._glx.glXDestroyContext(self._dpy, c_void_p(self._ctx))
            self._ctx = None


def available() -> bool:
    try:
        _load()
        shader_sources()
        return True
    except GLUnavailable:
        return False


def render_scene(*, width, height, scalar, normals, bmin, bmax, lut, view, proj, cam_pos, step_size, max_steps,
                 ref_step, ambient, diffuse, light_pos, light_target, threads=None, renderer=None):
    """Same keyword arguments as ``oracle.render_scene``; returns the RGBA8 frame ``(H, W, 4) uint8``
    (row 0 = bottom), i.e. ``np.frombuffer(VolumeRenderer.render())`` of the reference."""
    r = renderer or GLReference(width, height, threads=threads)
    if r.program is None:
        r.load_shaders()
    r.set_config(step_size, max_steps, ref_step)
    r.set_light(ambient, diffuse, light_pos, light_target)
    r.load_volume(scalar, normals, bmin, bmax)
    r.set_camera(view, proj, cam_pos)
    r.set_lut(lut)
    frame = np.frombuffer(r.render(), dtype=np.uint8).reshape(height, width, 4).copy()
    if renderer is None:
        r.close()
    return frame


def render(volume, camera, light, config, lut, width, height, threads=None):
    """Host objects in (the repo's mirror classes or the reference's own), frame out -- the call sequence of
    ``examples/benchmark.py:89-110``: construct, load_volume, set_camera, set_transfer_functions, render."""
    position, _ = camera.get_camera_vectors()
    return render_scene(
        width=width, height=height, scalar=volume.data, normals=volume.normals, bmin=volume.min_bounds,
        bmax=volume.max_bounds, lut=lut, view=camera.get_view_matrix(),
        proj=camera.get_projection_matrix(width / height), cam_pos=position, step_size=config.step_size,
        max_steps=config.max_steps, ref_step=config.reference_step_size, ambient=light.ambient_intensity,
        diffuse=light.diffuse_intensity, light_pos=light.position, light_target=light.target, threads=threads)
