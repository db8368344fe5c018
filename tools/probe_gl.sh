#!/bin/bash
# Can the reference's GL path (moderngl standalone context, manager.py:26) run on this box?  Records what the
# box has: GL/EGL libraries, glvnd vendor files, the driver capabilities the container was started with, and
# whether moderngl / glcontext import.  Output: gpurun_out/<tag>_gl_probe.txt (copied to profiles/).
TAG=${1:-r02}
OUT=gpurun_out/${TAG}_gl_probe.txt
mkdir -p gpurun_out
{
echo "## date: $(date -u +%FT%TZ)"
echo "## nvidia-smi"; nvidia-smi --query-gpu=name,driver_version --format=csv,noheader
echo "## env NVIDIA_DRIVER_CAPABILITIES=${NVIDIA_DRIVER_CAPABILITIES:-<unset>} NVIDIA_VISIBLE_DEVICES=${NVIDIA_VISIBLE_DEVICES:-<unset>}"
echo "## ldconfig -p | grep -Ei 'egl|opengl|glx|libGL|gbm|osmesa|glapi|vulkan'"
ldconfig -p | grep -Ei 'egl|opengl|glx|libGL|gbm|osmesa|glapi|vulkan' || echo "(none)"
echo "## ls libEGL* libnvidia-*gl* libGLX* libnvidia-egl*"
ls -la /usr/lib/x86_64-linux-gnu/libEGL* /usr/lib/x86_64-linux-gnu/libnvidia-*gl* /usr/lib/x86_64-linux-gnu/libGLX* \
       /usr/lib/x86_64-linux-gnu/libnvidia-egl* /usr/lib/x86_64-linux-gnu/libGLES* /usr/lib/x86_64-linux-gnu/libOpenGL* 2>&1 | grep -v "No such file" || true
echo "## find / -xdev -name 'libEGL*' -o -name 'libGLX_nvidia*' -o -name 'libnvidia-glcore*' -o -name 'libOSMesa*' -o -name 'libgbm*'"
find / -xdev \( -name 'libEGL*' -o -name 'libGLX_nvidia*' -o -name 'libnvidia-glcore*' -o -name 'libnvidia-eglcore*' -o -name 'libOSMesa*' -o -name 'libgbm*' -o -name 'libGL.so*' \) 2>/dev/null | head -40 || true
echo "## glvnd vendor dirs"
ls -la /usr/share/glvnd/egl_vendor.d /etc/glvnd/egl_vendor.d /usr/share/egl/egl_external_platform.d 2>&1 | head -20
echo "## libnvidia-* present"
ls /usr/lib/x86_64-linux-gnu/ | grep -i nvidia | head -60
echo "## /dev/dri /dev/nvidia*"
ls -la /dev/dri /dev/nvidia* 2>&1 | head -20
echo "## python imports"
python - <<'PY'
import importlib
for m in ("moderngl", "glcontext", "OpenGL", "vtk", "matplotlib", "PIL", "scipy", "cv2"):
    try:
        mod = importlib.import_module(m)
        print(m, "OK", getattr(mod, "__version__", ""))
    except Exception as e:
        print(m, "MISSING", type(e).__name__, e)
import ctypes, ctypes.util
for name in ("EGL", "GL", "OpenGL", "GLESv2", "OSMesa", "gbm"):
    print("find_library", name, "->", ctypes.util.find_library(name))
for path in ("libEGL.so.1", "libEGL_nvidia.so.0", "libGL.so.1", "libOpenGL.so.0"):
    try:
        ctypes.CDLL(path); print("dlopen", path, "OK")
    except OSError as e:
        print("dlopen", path, "FAILED", e)
PY
} > $OUT 2>&1
cat $OUT
