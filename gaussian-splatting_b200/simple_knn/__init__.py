"""Drop-in for the reference's absent ``simple_knn`` submodule: ``from simple_knn._C import distCUDA2``
(/root/reference/scene/gaussian_model.py:21), backed by libgs_b200.so (csrc/knn.cu)."""
