"""Manager / lifecycle loop — mirror of pkg/gpu/nvidia/gpumanager.go."""
from __future__ import annotations

import logging
import os
import queue
import signal
import threading
import time

from .. import device
from .._abi import GsbError
from . import const, nvidia
from .coredump import coredump
from .server import NewNvidiaDevicePlugin
from .watchers import IN_CREATE, newFSWatcher, newOSWatcher

log = logging.getLogger("gpushare.nvidia")


class sharedGPUManager:
    def __init__(self, enableMPS: bool, healthCheck: bool, queryKubelet: bool, kubeletClient,
                 pluginDir: str = const.DevicePluginPath, dumpDir: str = "/etc/kubernetes/", **plugin_kw):
        self.enableMPS, self.healthCheck, self.queryKubelet, self.kubeletClient = (enableMPS, healthCheck,
                                                                                 queryKubelet, kubeletClient)
        self.pluginDir, self.dumpDir, self.plugin_kw = pluginDir, dumpDir, plugin_kw
        self.devicePlugin = None

    def Run(self) -> None:  # gpumanager.go:33-111
        log.info("Loading NVML")
        try:
            device.init()  # nvml.Init()
        except GsbError as e:
            log.info("Failed to initialize NVML: %s.", e)
            log.info("If this is a GPU node, did you set the docker default runtime to `nvidia`?")
            threading.Event().wait()  # select {}: park forever, no crash loop (gpumanager.go:36-40)
        try:
            log.info("Fetching devices.")
            if nvidia.getDeviceCount() == 0:
                log.info("No devices found. Waiting indefinitely.")
                threading.Event().wait()  # gpumanager.go:44-47
            log.info("Starting FS watcher.")
            watcher = newFSWatcher(self.pluginDir)
            log.info("Starting OS watcher.")
            sigs = newOSWatcher(signal.SIGHUP, signal.SIGINT, signal.SIGTERM, signal.SIGQUIT)
            kubeletSock = os.path.join(self.pluginDir, "kubelet.sock")
            socket = os.path.join(self.pluginDir, os.path.basename(const.serverSock))
            restart = True
            try:
                while True:
                    if restart:
                        if self.devicePlugin is not None:
                            self.devicePlugin.Stop()
                        try:
                            self.devicePlugin = NewNvidiaDevicePlugin(self.enableMPS, self.healthCheck,
                                                                      self.queryKubelet, self.kubeletClient,
                                                                      socket=socket, **self.plugin_kw)
                        except Exception as e:  # noqa: BLE001
                            log.warning("Failed to get device plugin due to %s", e)
                            os._exit(1)  # gpumanager.go:73
                        try:
                            self.devicePlugin.Serve(kubeletSock)
                        except Exception as e:  # noqa: BLE001
                            log.warning("Failed to start device plugin due to %s", e)
                            os._exit(2)  # gpumanager.go:76
                        restart = False
                    # select over watcher.Events / watcher.Errors / sigs (gpumanager.go:82-107)
                    try:
                        name, mask = watcher.Events.get(timeout=0.1)
                        if name == kubeletSock and mask & IN_CREATE:
                            log.info("inotify: %s created, restarting.", kubeletSock)
                            restart = True
                        continue
                    except queue.Empty:
                        pass
                    try:
                        log.warning("inotify: %s", watcher.Errors.get_nowait())
                    except queue.Empty:
                        pass
                    try:
                        s = sigs.get_nowait()
                    except queue.Empty:
                        continue
                    if s == signal.SIGHUP:
                        log.info("Received SIGHUP, restarting.")
                        restart = True
                    elif s == signal.SIGQUIT:
                        log.info("generate core dump")
                        coredump(os.path.join(self.dumpDir, "go_" + time.strftime("%Y%m%d%H%M%S") + ".txt"))
                    else:
                        log.info('Received signal "%s", shutting down.', signal.Signals(s).name)
                        self.devicePlugin.Stop()
                        break
            finally:
                watcher.Close()
        finally:
            try:
                device.shutdown()
                log.info("Shutdown of NVML returned: <nil>")
            except GsbError as e:
                log.info("Shutdown of NVML returned: %s", e)


def NewSharedGPUManager(enableMPS: bool, healthCheck: bool, queryKubelet: bool, bp: str, client,
                        **kw) -> sharedGPUManager:  # gpumanager.go:23-31
    nvidia.metric = bp
    return sharedGPUManager(enableMPS, healthCheck, queryKubelet, client, **kw)
