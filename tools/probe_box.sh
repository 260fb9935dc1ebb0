#!/bin/bash
# One-off probe of the GPU box: is there a Bullet build or a software Vulkan ICD that could pin the two unpinned holes?
echo "== python modules"; python - <<'PY'
import importlib
for m in ("pybullet", "pybullet_data", "vulkan", "moderngl", "OpenGL", "glfw", "mujoco"):
    try:
        importlib.import_module(m); print(m, "present")
    except Exception as e:
        print(m, "absent:", type(e).__name__)
PY
echo "== libraries"; ldconfig -p | grep -i -E "bullet|vulkan|lvp|lavapipe|EGL|libGL|osmesa|swrast" || echo "(none in ldconfig)"
find / -xdev \( -iname "libBullet*" -o -iname "libLinearMath*" -o -iname "btBulletDynamicsCommon.h" -o -iname "*lvp_icd*" -o -iname "libvulkan*" -o -iname "vulkan.h" -o -iname "*nvidia_icd*.json" \) 2>/dev/null | head -40
ls /usr/share/vulkan/icd.d /etc/vulkan/icd.d 2>/dev/null
echo "== cpu"; nproc; lscpu | grep -E "Model name|NUMA|Socket" ; numactl --hardware 2>/dev/null | head -5
nvidia-smi topo -m 2>/dev/null | head -20
