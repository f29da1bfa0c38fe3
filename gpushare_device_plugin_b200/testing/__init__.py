"""Test doubles for the control plane either side of the path: a stateful mock kube-apiserver +
kubelet /pods/ endpoint, a fake kubelet Registration server, and synthetic cluster generators for
SURVEY.md §8(d) configs 4-5. Used by tests/ and by bench_allocate; never by the plugin itself."""
