"""`iopath.common.file_io.g_pathmgr` stand-in (local filesystem only) for the offline reference run."""
import os


class _LocalPathManager:
    def open(self, path, mode="r", **kw):
        return open(path, mode, **kw)

    def mkdirs(self, path):
        os.makedirs(path, exist_ok=True)

    def exists(self, path):
        return os.path.exists(path)

    def ls(self, path):
        return os.listdir(path)

    def isfile(self, path):
        return os.path.isfile(path)


g_pathmgr = _LocalPathManager()
PathManager = g_pathmgr
