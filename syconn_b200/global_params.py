"""Minimal stand-in for ``syconn.global_params``: only the keys the hot path reads
(syconn/handler/config.yml:148-150; read at syconn/extraction/find_object_properties.py:471)."""
config = {"cell_objects": {"cs_filtersize": [13, 13, 7], "cs_dilation": 2}}
