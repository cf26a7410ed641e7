/* Non-Python host for libpcb_b200.so: evaluate a `.pcb` file at a batch of points.
 *
 *   pcb_eval model.pcb points.f64 out.f64
 *
 * points.f64: N x D doubles, row-major (D is read from the model); out.f64: N doubles.
 * The counterpart of the reference's examples/binary_reader/reader.c, with the evaluation on the
 * GPU: pcb_plan_from_file (include/pcb_b200.h) parses the file and builds the device plan,
 * pcb_plan_eval runs one launch on the given stream.
 *
 * Build:  gcc -O2 -I include -I /usr/local/cuda/include examples/pcb_eval.c \
 *             -L pychebyshev_b200 -lpcb_b200 -L /usr/local/cuda/lib64 -lcudart \
 *             -Wl,-rpath,'$ORIGIN/../pychebyshev_b200' -o examples/pcb_eval
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "pcb_b200.h"

static int die(const char *what) {
    fprintf(stderr, "pcb_eval: %s: %s\n", what, pcb_last_error());
    return 1;
}

int main(int argc, char **argv) {
    if (argc != 4) {
        fprintf(stderr, "usage: %s model.pcb points.f64 out.f64\n", argv[0]);
        return 2;
    }
    void *plan = NULL;
    int kind = 0, D = 0;
    if (pcb_plan_from_file(0, argv[1], &plan, &kind, &D) != PCB_OK) return die("pcb_plan_from_file");

    FILE *f = fopen(argv[2], "rb");
    if (!f) {
        perror(argv[2]);
        return 1;
    }
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    const long long N = bytes / (8LL * D);
    double *h_pts = (double *)malloc((size_t)bytes), *h_out = (double *)malloc((size_t)N * 8);
    if (fread(h_pts, 1, (size_t)bytes, f) != (size_t)bytes) {
        fprintf(stderr, "short read on %s\n", argv[2]);
        return 1;
    }
    fclose(f);

    double *d_pts = NULL, *d_out = NULL;
    cudaStream_t stream;
    if (cudaStreamCreate(&stream) != cudaSuccess || cudaMalloc((void **)&d_pts, (size_t)bytes) != cudaSuccess ||
        cudaMalloc((void **)&d_out, (size_t)N * 8) != cudaSuccess) {
        fprintf(stderr, "pcb_eval: CUDA allocation failed\n");
        return 1;
    }
    cudaMemcpyAsync(d_pts, h_pts, (size_t)bytes, cudaMemcpyHostToDevice, stream);
    if (pcb_plan_eval(plan, d_pts, N, d_out, stream) != PCB_OK) return die("pcb_plan_eval");
    cudaMemcpyAsync(h_out, d_out, (size_t)N * 8, cudaMemcpyDeviceToHost, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess) {
        fprintf(stderr, "pcb_eval: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    f = fopen(argv[3], "wb");
    if (!f || fwrite(h_out, 8, (size_t)N, f) != (size_t)N) {
        perror(argv[3]);
        return 1;
    }
    fclose(f);
    printf("%s: %s, %d dims, %lld points, %llu kernel launch(es)\n", argv[1],
           kind == 1 ? "ChebyshevApproximation" : "ChebyshevSpline", D, N,
           (unsigned long long)pcb_launch_count());
    pcb_plan_destroy(plan);
    cudaFree(d_pts);
    cudaFree(d_out);
    cudaStreamDestroy(stream);
    free(h_pts);
    free(h_out);
    return 0;
}
