#!/bin/bash
# Second-stage probe: what do the NVIDIA EGL/GL vendor libraries on the box export and depend on?
OUT=gpurun_out/r02_gl_probe2.txt
mkdir -p gpurun_out
{
ls -laL /usr/lib/libEGL_nvidia* /usr/lib/libGLX_nvidia* /usr/lib/libnvidia-eglcore* /usr/lib/libnvidia-glcore* /usr/lib/libnvidia-glsi* /usr/lib/libnvidia-tls* /usr/lib/libnvidia-gpucomp* /usr/lib/libnvidia-glvkspirv* 2>&1
echo "## all nvidia libs anywhere"
find / -xdev \( -name 'libnvidia-*' -o -name 'libGLdispatch*' -o -name 'libcuda.so*' -o -name 'libnvidia-ml*' \) 2>/dev/null | head -60
echo "## readelf -d libEGL_nvidia.so.0"
readelf -d /usr/lib/libEGL_nvidia.so.0 | grep -E 'NEEDED|SONAME|RPATH|RUNPATH'
echo "## exported symbols libEGL_nvidia.so.0"
nm -D --defined-only /usr/lib/libEGL_nvidia.so.0 | head -80
echo "## readelf -d libnvidia-eglcore.so"
readelf -d /usr/lib/libnvidia-eglcore.so | grep -E 'NEEDED|SONAME'
echo "## readelf -d libGLX_nvidia.so.0"
readelf -d /usr/lib/libGLX_nvidia.so.0 | grep -E 'NEEDED|SONAME'
echo "## exported symbols libGLX_nvidia.so.0 (first 40)"
nm -D --defined-only /usr/lib/libGLX_nvidia.so.0 | head -40
echo "## strings: version"
strings /usr/lib/libEGL_nvidia.so.0 | grep -E '^[0-9]{3}\.[0-9]+' | head -5
echo "## ld.so.conf"
cat /etc/ld.so.conf /etc/ld.so.conf.d/* 2>/dev/null | head -20
} > $OUT 2>&1
cat $OUT
