from .launch import map_launcher_env  # noqa: F401
