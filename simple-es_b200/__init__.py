"""Product directory of the simple-es B200 engine; import it as ``simple_es_b200`` (see ../simple_es_b200/__init__.py)."""
