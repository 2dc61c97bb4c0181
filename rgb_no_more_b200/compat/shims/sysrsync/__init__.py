"""`sysrsync.run` as /datasets.py of the reference uses it (datasets.py:101,125): local copy through the rsync binary when
present, shutil otherwise.  Dataset staging is outside the hot path."""
import os
import shutil
import subprocess


def run(source, destination, options=None, sync_source_contents=True, **kwargs):
    if shutil.which("rsync"):
        src = source + ("/" if sync_source_contents and os.path.isdir(source) and not source.endswith("/") else "")
        return subprocess.run(["rsync", *(options or ["-a"]), src, destination], check=True)
    if os.path.isdir(source):
        return shutil.copytree(source, destination, dirs_exist_ok=True)
    return shutil.copy2(source, destination)
