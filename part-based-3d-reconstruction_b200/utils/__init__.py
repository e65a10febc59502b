"""Reference-compatible `utils` package (same module and function names as the reference's
utils/ for the carving and camera-scoring path).  Unlike the reference's utils/__init__.py:1-9 this
does not star-import every module, so importing it pulls in no plotting dependencies."""
