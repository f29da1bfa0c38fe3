"""The wire/env/annotation contract of the gpushare device plugin, value for value.

Nothing here is a design choice of this repository: every string is something another component
(kubelet, the gpushare scheduler extender, tenant containers, `kubectl inspect gpushare`) matches
byte for byte. Grouped by WHO reads it; the reference defines them in pkg/gpu/nvidia/const.go:11-35 and
vendor/k8s.io/kubernetes/pkg/kubelet/apis/deviceplugin/v1beta1/constants.go:19-36.
"""

# ---- read by the kubelet's device manager --------------------------------------------------------
Version = "v1beta1"                                   # RegisterRequest.version the kubelet accepts
DevicePluginPath = "/var/lib/kubelet/device-plugins/"  # directory the kubelet watches for plugin sockets
KubeletSocket = DevicePluginPath + "kubelet.sock"      # Registration service (we are its client)
serverSock = DevicePluginPath + "aliyungpushare.sock"  # our DevicePlugin service (RegisterRequest.endpoint is its basename)
resourceName = "aliyun.com/gpu-mem"                    # extended resource: one unit = one GiB (or MiB) slice of one GPU
Healthy, Unhealthy = "Healthy", "Unhealthy"            # the only two Device.health values

# ---- read by the scheduler extender / written into node status -------------------------------------
resourceCount = "aliyun.com/gpu-count"                 # node capacity+allocatable: number of physical GPUs
EnvNodeLabelForDisableCGPU = "cgpu.disable.isolation"  # node label; "true" => containers get CGPU_DISABLE=true

# ---- pod annotations exchanged with the scheduler extender ------------------------------------------
EnvResourceIndex = "ALIYUN_COM_GPU_MEM_IDX"            # which GPU (by /dev/nvidia MINOR) the extender picked
EnvAssignedFlag = "ALIYUN_COM_GPU_MEM_ASSIGNED"        # "false" until Allocate claims the pod, then "true"
EnvResourceAssumeTime = "ALIYUN_COM_GPU_MEM_ASSUME_TIME"  # ns timestamp; oldest unassigned pod of the right size wins
EnvResourceAssignTime = "ALIYUN_COM_GPU_MEM_ASSIGN_TIME"  # defined by the reference, never written by it

# ---- environment injected into tenant containers by Allocate ----------------------------------------
envNVGPU = "NVIDIA_VISIBLE_DEVICES"                    # the GPU index (or UUID in the single-GPU shortcut, or the poison text)
EnvResourceByPod = "ALIYUN_COM_GPU_MEM_POD"            # slices requested by the whole pod
EnvResourceByContainer = "ALIYUN_COM_GPU_MEM_CONTAINER"  # slices requested by this container
EnvResourceByDev = "ALIYUN_COM_GPU_MEM_DEV"            # slices one GPU has in total (179 on a B200)

# ---- apiserver error text that earns the PATCH exactly one retry (allocate.go:138-144) ---------------
OptimisticLockErrorMsg = ("the object has been modified; please apply your changes to the latest version and "
                          "try again")

# ---- --memory-unit values ---------------------------------------------------------------------------
GiBPrefix, MiBPrefix = "GiB", "MiB"
