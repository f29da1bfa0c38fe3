"""Constants and the env/annotation contract (pkg/gpu/nvidia/const.go:11-35; v1beta1/constants.go)."""

resourceName = "aliyun.com/gpu-mem"
resourceCount = "aliyun.com/gpu-count"
DevicePluginPath = "/var/lib/kubelet/device-plugins/"      # v1beta1/constants.go:30
KubeletSocket = DevicePluginPath + "kubelet.sock"          # v1beta1/constants.go:32
serverSock = DevicePluginPath + "aliyungpushare.sock"      # const.go:13
Version = "v1beta1"
Healthy = "Healthy"
Unhealthy = "Unhealthy"

OptimisticLockErrorMsg = ("the object has been modified; please apply your changes to the latest version and "
                          "try again")

envNVGPU = "NVIDIA_VISIBLE_DEVICES"
EnvResourceIndex = "ALIYUN_COM_GPU_MEM_IDX"
EnvResourceByPod = "ALIYUN_COM_GPU_MEM_POD"
EnvResourceByContainer = "ALIYUN_COM_GPU_MEM_CONTAINER"
EnvResourceByDev = "ALIYUN_COM_GPU_MEM_DEV"
EnvAssignedFlag = "ALIYUN_COM_GPU_MEM_ASSIGNED"
EnvResourceAssumeTime = "ALIYUN_COM_GPU_MEM_ASSUME_TIME"
EnvResourceAssignTime = "ALIYUN_COM_GPU_MEM_ASSIGN_TIME"
EnvNodeLabelForDisableCGPU = "cgpu.disable.isolation"

GiBPrefix = "GiB"
MiBPrefix = "MiB"
