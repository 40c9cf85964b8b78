"""CPU oracle for the YoloSharp forward/NMS hot path.

TEST INFRASTRUCTURE ONLY.  This package is a PyTorch-CPU (fp32) restatement of
the reference's op sequence (IntptrMax/YoloSharp @ 16dc3cd).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product package ``yolosharp_b200`` never
does.

PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer
values for this path, and its C#/TorchSharp code cannot be
executed in this environment (no .NET toolchain).  The arithmetic of the
reference lives in libtorch (TorchSharp 0.105.2 -> libtorch 2.5.1 / 2.7.1);
this oracle calls the same operator library family through PyTorch 2.11 CPU.
The only external anchors are the shipped checkpoints + test images (e.g.
``Yolov8n.bin`` on ``bus.jpg`` -> bus 0.896 + 3 persons), checked in
``tests/test_oracle.py`` when ``/root/reference`` is present.
"""
