/*
 * ref_launcher.c — 20-line front door of oracle/_ref/libref_inventory.so (test infrastructure).
 *
 * The reference links its cgo object with -Wl,--unresolved-symbols=ignore-in-object-files
 * (bindings.go:5) and relies on lazy PLT binding: nvml* symbols resolve at first call, after
 * nvmlInit_dl() has dlopen'ed libnvidia-ml.so.1 RTLD_GLOBAL. A plain C *executable* linked that way
 * gets null call targets instead of PLT slots, so the restatement + the reference's nvml_dl.c are
 * built as a lazily-bound shared object (which keeps the reference's mechanism intact) and this
 * launcher loads it.
 */
#include <dlfcn.h>
#include <libgen.h>
#include <limits.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

int main(int argc, char **argv) {
  char self[PATH_MAX], lib[PATH_MAX + 32];
  ssize_t n = readlink("/proc/self/exe", self, sizeof self - 1);
  if (n <= 0) return 3;
  self[n] = 0;
  snprintf(lib, sizeof lib, "%s/libref_inventory.so", dirname(self));
  void *h = dlopen(lib, RTLD_LAZY | RTLD_GLOBAL);
  if (!h) {
    fprintf(stderr, "%s\n", dlerror());
    return 3;
  }
  int (*ref_main)(int, char **) = (int (*)(int, char **))dlsym(h, "ref_main");
  return ref_main ? ref_main(argc, argv) : 3;
}
