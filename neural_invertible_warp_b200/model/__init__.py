"""Drop-in ``Graph`` / ``NeRF`` classes, one module per reference model file
(``model/nerf.py``, ``barf.py``, ``nerf_inn_llff.py``, ``barf_inn_llff.py``, ``nerf_inn_dtu.py``,
``barf_inn_dtu.py``).  Import a module by the reference's model name:

    importlib.import_module("neural_invertible_warp_b200.model." + opt.model).Graph(opt)
"""
