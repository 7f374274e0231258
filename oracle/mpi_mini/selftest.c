/* TEST INFRASTRUCTURE ONLY -- self-test of oracle/mpi_mini (tests/test_reference_host_multirank_cpu.py) */
#include <stdio.h>
#include <stdlib.h>
#include "mpi.h"
int main(int argc, char **argv)
{
	MPI_Init(&argc, &argv);
	int r, n; MPI_Comm_rank(0, &r); MPI_Comm_size(0, &n);
	int L = (r + n - 1) % n, R = (r + 1) % n;
	/* big ring sendrecv: 8 MB each way, everybody sends first */
	long cnt = 1 << 20; double *a = malloc(cnt * 8), *b = malloc(cnt * 8);
	for (long i = 0; i < cnt; i++) a[i] = r * 1e6 + i;
	MPI_Sendrecv(a, cnt, MPI_DOUBLE, R, 7, b, cnt, MPI_DOUBLE, L, 7, 0, MPI_STATUS_IGNORE);
	int ok = b[5] == L * 1e6 + 5 && b[cnt - 1] == L * 1e6 + cnt - 1;
	/* same tag, several messages, non-overtaking; irecv posted in reverse of arrival for different tags */
	MPI_Request sq[6], rq[6]; double s[6], t[6];
	for (int i = 0; i < 6; i++) { s[i] = 100 * r + i; MPI_Isend(&s[i], 1, MPI_DOUBLE, R, i % 2, 0, &sq[i]); }
	for (int i = 5; i >= 0; i--) MPI_Irecv(&t[i], 1, MPI_DOUBLE, L, i % 2, 0, &rq[i]);
	MPI_Waitall(6, rq, MPI_STATUSES_IGNORE);
	/* posting order 5,3,1 on tag 1 receives messages 1,3,5 in order -> t[5]=s1,t[3]=s3,t[1]=s5 of rank L */
	ok = ok && t[5] == 100 * L + 1 && t[3] == 100 * L + 3 && t[1] == 100 * L + 5 && t[4] == 100 * L + 0 && t[2] == 100 * L + 2 && t[0] == 100 * L + 4;
	double x = r + 1.5, y = 0; MPI_Allreduce(&x, &y, 1, MPI_DOUBLE, MPI_SUM, 0);
	double want = 0; for (int i = 0; i < n; i++) want += i + 1.5;
	ok = ok && y == want;
	int m = r * 3, mm = 0; MPI_Allreduce(&m, &mm, 1, MPI_INT, MPI_MAX, 0); ok = ok && mm == 3 * (n - 1);
	int v[4] = { r, r, r, r }; MPI_Bcast(v, 4, MPI_INT, 0, 0); ok = ok && v[3] == 0;
	char c[3] = { 'a' + r, 0, 0 }; MPI_Bcast(c, 3, MPI_CHAR, n - 1, 0); ok = ok && c[0] == 'a' + n - 1;
	MPI_Barrier(0);
	if (r == 0) { double z = 3.25; for (int p = 1; p < n; p++) MPI_Send(&z, 1, MPI_DOUBLE, p, p, 0); }
	else { double z = 0; MPI_Recv(&z, 1, MPI_DOUBLE, 0, r, 0, MPI_STATUS_IGNORE); ok = ok && z == 3.25; }
	printf("rank %d of %d: %s\n", r, n, ok ? "OK" : "FAILED");
	MPI_Finalize();
	return ok ? 0 : 1;
}
