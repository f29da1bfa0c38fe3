/* A plain-C99 caller of libgpushare_b200.so — what a cgo binding sees. Host-only entry points:
 * it runs on a box without a GPU and prints one line per call; tests/test_abi.py compares the lines with the
 * golden vectors and the oracle. Built with: gcc -std=c99 -pedantic -Wall -Wextra -Werror */
#include <stdio.h>
#include <string.h>

#include "gpushare_b200.h"

static void hex(const uint8_t *p, size_t n) {
  size_t i;
  for (i = 0; i < n; i++) printf("%02x", p[i]);
  printf("\n");
}

int main(void) {
  static const char *uuids[2] = {"GPU-fef8089b-4820-abfc-e83e-94318197576e", "GPU-fef8089c-4820-abfc-e83e-94318197576f"};
  static const uint32_t minors[2] = {3, 1};
  static uint8_t big[1 << 16];
  char id[128], real[128], patch[256];
  uint8_t bits[64], resp[4096];
  /* AllocateRequest{container_requests:[{devicesIDs:["a","b","c","d"]}]} */
  static const uint8_t req[] = {0x0a, 0x0c, 0x0a, 0x01, 'a', 0x0a, 0x01, 'b', 0x0a, 0x01, 'c', 0x0a, 0x01, 'd'};
  gsb_pod pods[2];
  gsb_allocate_ctx ctx;
  size_t n = 0;
  int32_t pod_index = -1;
  uint32_t pod_req = 0;
  int rc;
  int64_t len;

  printf("abi %d\n", gsb_abi_version());
  printf("slices %u %u %u\n", gsb_slices(183359, 1), gsb_slices(183359, 0), gsb_slices(1023, 1));
  rc = gsb_fake_device_id(uuids[0], 178, id, sizeof id);
  printf("fake %d %s\n", rc, id);
  rc = gsb_real_device_id(id, real, sizeof real);
  printf("real %d %s\n", rc, real);
  printf("small-buffer %s\n", gsb_strerror(gsb_fake_device_id(uuids[0], 178, id, 8)));
  printf("xid %d %d %d %d %d\n", gsb_xid_is_benign(31), gsb_xid_is_benign(43), gsb_xid_is_benign(45), gsb_xid_is_benign(48),
         gsb_xid_is_benign(79));

  memset(bits, 0, sizeof bits);
  bits[0] = 0x02; /* fake device 1 of GPU 0 is Unhealthy */
  len = gsb_encode_list_and_watch(uuids, 2, 3, bits, big, sizeof big);
  printf("lw %ld ", (long)len);
  hex(big, len > 0 ? (size_t)len : 0);
  len = gsb_encode_list_and_watch(uuids, 1, 179, NULL, big, sizeof big);
  printf("lw179 %ld\n", (long)len);
  len = gsb_encode_register_request("v1beta1", "aliyungpushare.sock", "aliyun.com/gpu-mem", big, sizeof big);
  printf("register %ld ", (long)len);
  hex(big, len > 0 ? (size_t)len : 0);

  memset(pods, 0, sizeof pods);
  pods[0].name = "pod-00"; pods[0].ns = "default"; pods[0].uid = "uid-0";
  pods[0].gpu_mem_limit = 4; pods[0].assume_time = 20; pods[0].gpu_idx = 3;
  pods[0].has_assume_time = 1; pods[0].has_assigned = 1; pods[0].assigned_is_false = 1; pods[0].on_node = 1;
  pods[1] = pods[0];
  pods[1].name = "pod-01"; pods[1].uid = "uid-1"; pods[1].assume_time = 10; pods[1].gpu_idx = 1;
  memset(&ctx, 0, sizeof ctx);
  ctx.uuids = uuids; ctx.minors = minors; ctx.n_gpus = 2; ctx.slices = 179; ctx.unit_gib = 1;
  rc = gsb_allocate(&ctx, pods, 2, req, sizeof req, resp, sizeof resp, &n, &pod_index, &pod_req);
  printf("allocate %d pod %d req %u ", rc, (int)pod_index, pod_req);
  hex(resp, n);
  pods[1].assigned_is_false = 0; /* claimed: the older pod is gone, the next request gets pod-00 */
  rc = gsb_allocate(&ctx, pods, 2, req, sizeof req, resp, sizeof resp, &n, &pod_index, &pod_req);
  printf("allocate %d pod %d req %u ", rc, (int)pod_index, pod_req);
  hex(resp, n);
  rc = gsb_allocate_err_response(&ctx, req, sizeof req, resp, sizeof resp, &n);
  printf("err %d ", rc);
  hex(resp, n);
  rc = gsb_patch_assigned_body(1700000000000000000ull, patch, sizeof patch);
  printf("patch %d %s\n", rc, patch);
  /* no driver on this box: the device entry points must say so, not pretend */
  rc = gsb_device_count(&pod_req);
  printf("device_count %s\n", rc == GSB_OK ? "ok" : gsb_strerror(rc));
  return 0;
}
