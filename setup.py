"""Packaging (reference: python/setup.py.in -- package ``edl``, console script ``edlrun``).
The native extension is built in-tree by ``python -m edl_b200.build_ext`` (nvcc, sm_100a)."""
from setuptools import find_packages, setup

setup(
    name="paddle_edl_b200",
    version="0.1.0",
    description="Blackwell-native elastic deep-learning engine with the paddle_edl API",
    packages=find_packages(include=["edl_b200*", "paddle_edl*", "edl*"]),
    package_data={"edl_b200": ["_C*.so", "csrc/*"]},
    python_requires=">=3.9",
    install_requires=["torch", "grpcio", "protobuf", "msgpack", "psutil", "numpy"],
    entry_points={"console_scripts": ["edlrun = edl_b200.collective.launch:run_commandline"]},
)
