/* TEST INFRASTRUCTURE ONLY -- see mpi.h.  Full mesh of Unix-domain stream sockets between N local processes; every message
 * is a frame {tag, bytes} + payload; per-peer FIFO of unexpected frames gives MPI's non-overtaking matching on (source, tag).
 * Sends are eager and never deadlock: while a socket would block for writing, incoming frames of every peer are drained.
 * Reductions are accumulated on rank 0 in rank order, so sums are deterministic and equal on all ranks. */
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>
#include "mpi.h"

#define MAXR 64
typedef struct msg { int tag; long bytes; char *data; struct msg *next; } msg;
typedef struct { int fd; char *buf; long have, cap; msg *head, *tail; } peer;
static peer P[MAXR];
static int g_rank = 0, g_size = 1, g_init = 0;
typedef struct { int active /* 1 posted, 2 matched */, src, tag; long bytes, seq; void *buf; } req;
static req R[4096];
static long g_seq = 0;
enum { TAG_BARRIER = -10, TAG_BCAST = -11, TAG_REDUCE = -12 };

static void die(const char *what) { fprintf(stderr, "mpi_mini[%d]: %s (%s)\n", g_rank, what, strerror(errno)); exit(1); }
static long dtsize(MPI_Datatype t) { return t == MPI_DOUBLE ? 8 : t == MPI_FLOAT ? 4 : t == MPI_INT ? 4 : 1; }

static int parse_frames(peer *q)
{	/* cut complete frames out of the byte buffer into the FIFO */
	long off = 0; int n = 0;
	const long hdr = (long) (sizeof(int) + sizeof(long));
	while (q->have - off >= hdr) {
		int tag; long bytes;
		memcpy(&tag, q->buf + off, sizeof(int)); memcpy(&bytes, q->buf + off + sizeof(int), sizeof(long));
		if (q->have - off < hdr + bytes) break;
		msg *m = (msg *) malloc(sizeof(msg));
		m->tag = tag; m->bytes = bytes; m->data = (char *) malloc(bytes ? bytes : 1); m->next = NULL;
		memcpy(m->data, q->buf + off + hdr, bytes);
		if (q->tail) q->tail->next = m; else q->head = m;
		q->tail = m;
		off += hdr + bytes; n++;
	}
	if (off) { memmove(q->buf, q->buf + off, q->have - off); q->have -= off; }
	return n;
}

/* read from peer p: block = 0 takes what is there, block = 1 returns once at least one new frame is complete */
static int drain(int p, int block)
{
	peer *q = &P[p];
	int got = 0;
	for (;;) {
		if (q->cap - q->have < 65536) { q->cap = q->cap ? 2 * q->cap : (1 << 20); q->buf = (char *) realloc(q->buf, q->cap); }
		ssize_t n = read(q->fd, q->buf + q->have, q->cap - q->have);
		if (n > 0) { q->have += n; got += parse_frames(q); if (block && got) return got; continue; }
		if (n == 0) {
			if (block) { fprintf(stderr, "mpi_mini[%d]: rank %d closed the connection\n", g_rank, p); exit(1); }
			return got;
		}
		if (errno == EINTR) continue;
		if (errno == EAGAIN || errno == EWOULDBLOCK) {
			if (!block) return got;
			struct pollfd pf = { q->fd, POLLIN, 0 };
			if (poll(&pf, 1, -1) < 0 && errno != EINTR) die("poll");
			continue;
		}
		die("read");
	}
}

static void send_bytes(int p, int tag, const void *buf, long bytes)
{
	if (p == g_rank) {            /* self-send: straight into the own queue */
		peer *q = &P[p];
		msg *m = (msg *) malloc(sizeof(msg));
		m->tag = tag; m->bytes = bytes; m->data = (char *) malloc(bytes ? bytes : 1); m->next = NULL; memcpy(m->data, buf, bytes);
		if (q->tail) q->tail->next = m; else q->head = m;
		q->tail = m; return;
	}
	long total = sizeof(int) + sizeof(long) + bytes, done = 0;
	char *frame = (char *) malloc(total);
	memcpy(frame, &tag, sizeof(int)); memcpy(frame + sizeof(int), &bytes, sizeof(long)); memcpy(frame + sizeof(int) + sizeof(long), buf, bytes);
	while (done < total) {
		ssize_t n = write(P[p].fd, frame + done, total - done);
		if (n > 0) { done += n; continue; }
		if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
			for (int r = 0; r < g_size; r++) if (r != g_rank) drain(r, 0);        /* keep everybody's pipes moving */
			struct pollfd pf = { P[p].fd, POLLOUT, 0 }; poll(&pf, 1, 10);
			continue;
		}
		if (n < 0 && errno == EINTR) continue;
		die("write");
	}
	free(frame);
}

static void recv_bytes(int p, int tag, void *buf, long bytes)
{
	for (;;) {
		msg *prev = NULL;
		for (msg *m = P[p].head; m; prev = m, m = m->next)
			if (m->tag == tag) {
				if (m->bytes != bytes) { fprintf(stderr, "mpi_mini[%d]: size mismatch from %d tag %d: %ld vs %ld\n", g_rank, p, tag, m->bytes, bytes); exit(1); }
				memcpy(buf, m->data, bytes);
				if (prev) prev->next = m->next; else P[p].head = m->next;
				if (P[p].tail == m) P[p].tail = prev;
				free(m->data); free(m); return;
			}
		if (p == g_rank) { fprintf(stderr, "mpi_mini[%d]: self-receive with nothing sent (tag %d)\n", g_rank, tag); exit(1); }
		drain(p, 1);
	}
}

int MPI_Init(int *argc, char ***argv)
{
	(void) argc; (void) argv;
	if (g_init) return 0;
	g_init = 1;
	const char *r = getenv("MINI_MPI_RANK"), *s = getenv("MINI_MPI_SIZE"), *d = getenv("MINI_MPI_DIR");
	if (!r || !s || !d) { g_rank = 0; g_size = 1; return 0; }              /* not under the launcher: a single rank */
	g_rank = atoi(r); g_size = atoi(s);
	if (g_size > MAXR) die("too many ranks");
	struct sockaddr_un a; memset(&a, 0, sizeof(a)); a.sun_family = AF_UNIX;
	int lfd = socket(AF_UNIX, SOCK_STREAM, 0);
	snprintf(a.sun_path, sizeof(a.sun_path), "%s/r%d", d, g_rank);
	unlink(a.sun_path);
	if (bind(lfd, (struct sockaddr *) &a, sizeof(a)) < 0 || listen(lfd, MAXR) < 0) die("bind/listen");
	for (int p = 0; p < g_rank; p++) {                                      /* connect to the lower ranks ... */
		int fd = socket(AF_UNIX, SOCK_STREAM, 0);
		snprintf(a.sun_path, sizeof(a.sun_path), "%s/r%d", d, p);
		int tries = 0;
		while (connect(fd, (struct sockaddr *) &a, sizeof(a)) < 0) { if (++tries > 20000) die("connect"); usleep(1000); }
		if (write(fd, &g_rank, sizeof(int)) != sizeof(int)) die("hello");
		P[p].fd = fd;
	}
	for (int k = g_rank + 1; k < g_size; k++) {                            /* ... and accept the higher ones */
		int fd = accept(lfd, NULL, NULL), who = -1;
		if (fd < 0 || read(fd, &who, sizeof(int)) != sizeof(int) || who <= g_rank || who >= g_size) die("accept");
		P[who].fd = fd;
	}
	close(lfd);
	for (int p = 0; p < g_size; p++) if (p != g_rank) fcntl(P[p].fd, F_SETFL, fcntl(P[p].fd, F_GETFL) | O_NONBLOCK);
	return 0;
}
int MPI_Finalize(void) { if (g_size > 1) MPI_Barrier(0); for (int p = 0; p < g_size; p++) if (p != g_rank && P[p].fd) close(P[p].fd); return 0; }
int MPI_Abort(MPI_Comm c, int e) { (void) c; fflush(stdout); exit(e ? e : 1); }
int MPI_Comm_rank(MPI_Comm c, int *r) { (void) c; *r = g_rank; return 0; }
int MPI_Comm_size(MPI_Comm c, int *n) { (void) c; *n = g_size; return 0; }
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *o) { (void) c; (void) color; (void) key; *o = 0; return 0; }   /* NREPLICAS = 1 */
int MPI_Get_processor_name(char *n, int *l) { strcpy(n, "localhost"); *l = 9; return 0; }

/* MPI matches receives in the order they were POSTED: before anything receives on (src, tag), the nonblocking receives
 * posted earlier on the same (src, tag) take their messages */
static void match_posted_before(int src, int tag, long seq)
{
	for (;;) {
		int first = -1;
		for (int i = 0; i < 4096; i++)
			if (R[i].active == 1 && R[i].src == src && R[i].tag == tag && R[i].seq < seq && (first < 0 || R[i].seq < R[first].seq)) first = i;
		if (first < 0) return;
		recv_bytes(src, tag, R[first].buf, R[first].bytes);
		R[first].active = 2;
	}
}
int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) { (void) c; send_bytes(dst, tag, b, n * dtsize(t)); return 0; }
int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *s)
{ (void) c; (void) s; match_posted_before(src, tag, g_seq + 1); recv_bytes(src, tag, b, n * dtsize(t)); return 0; }
int MPI_Sendrecv(const void *sb, int sn, MPI_Datatype st, int dst, int stag, void *rb, int rn, MPI_Datatype rt, int src, int rtag,
								 MPI_Comm c, MPI_Status *s)
{
	(void) c; (void) s;
	send_bytes(dst, stag, sb, sn * dtsize(st));
	match_posted_before(src, rtag, g_seq + 1); recv_bytes(src, rtag, rb, rn * dtsize(rt));
	return 0;
}
int MPI_Isend(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *rq)
{ (void) c; send_bytes(dst, tag, b, n * dtsize(t)); *rq = -1; return 0; }          /* eager: complete on return */
int MPI_Irecv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *rq)
{
	(void) c;
	for (int i = 0; i < 4096; i++)
		if (!R[i].active) { R[i].active = 1; R[i].src = src; R[i].tag = tag; R[i].bytes = n * dtsize(t); R[i].buf = b; R[i].seq = ++g_seq; *rq = i; return 0; }
	die("out of requests"); return 1;
}
int MPI_Wait(MPI_Request *rq, MPI_Status *s)
{
	(void) s;
	if (*rq >= 0 && R[*rq].active) {
		req *q = &R[*rq];
		if (q->active == 1) { match_posted_before(q->src, q->tag, q->seq); recv_bytes(q->src, q->tag, q->buf, q->bytes); }
		q->active = 0;
	}
	*rq = -1; return 0;
}
int MPI_Waitall(int n, MPI_Request *rq, MPI_Status *s) { for (int i = 0; i < n; i++) MPI_Wait(&rq[i], s); return 0; }

int MPI_Barrier(MPI_Comm c)
{
	(void) c; char z = 0;
	if (g_size == 1) return 0;
	if (g_rank == 0) { for (int p = 1; p < g_size; p++) recv_bytes(p, TAG_BARRIER, &z, 1); for (int p = 1; p < g_size; p++) send_bytes(p, TAG_BARRIER, &z, 1); }
	else { send_bytes(0, TAG_BARRIER, &z, 1); recv_bytes(0, TAG_BARRIER, &z, 1); }
	return 0;
}
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{
	(void) c;
	if (g_size == 1) return 0;
	if (g_rank == root) { for (int p = 0; p < g_size; p++) if (p != root) send_bytes(p, TAG_BCAST, b, n * dtsize(t)); }
	else recv_bytes(root, TAG_BCAST, b, n * dtsize(t));
	return 0;
}
static void combine(void *acc, const void *in, int n, MPI_Datatype t, MPI_Op op)
{
	for (int i = 0; i < n; i++) {
		if (t == MPI_DOUBLE) { double *a = (double *) acc; const double *b = (const double *) in; a[i] = op == MPI_SUM ? a[i] + b[i] : (a[i] > b[i] ? a[i] : b[i]); }
		else if (t == MPI_FLOAT) { float *a = (float *) acc; const float *b = (const float *) in; a[i] = op == MPI_SUM ? a[i] + b[i] : (a[i] > b[i] ? a[i] : b[i]); }
		else if (t == MPI_INT) { int *a = (int *) acc; const int *b = (const int *) in; a[i] = op == MPI_SUM ? a[i] + b[i] : (a[i] > b[i] ? a[i] : b[i]); }
		else { fprintf(stderr, "mpi_mini: reduction on an unsupported type\n"); exit(1); }
	}
}
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{
	(void) c; long bytes = n * dtsize(t);
	if (g_rank != root) { send_bytes(root, TAG_REDUCE, s, bytes); return 0; }
	char *tmp = (char *) malloc(bytes), *acc = (char *) malloc(bytes);
	for (int p = 0; p < g_size; p++) {                                     /* rank order: deterministic */
		if (p == root) memcpy(tmp, s, bytes); else recv_bytes(p, TAG_REDUCE, tmp, bytes);
		if (p == 0) memcpy(acc, tmp, bytes); else combine(acc, tmp, n, t, op);
	}
	memcpy(r, acc, bytes); free(tmp); free(acc);
	return 0;
}
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
	if (g_size == 1) { memcpy(r, s, n * dtsize(t)); return 0; }
	MPI_Reduce(s, r, n, t, op, 0, c);
	return MPI_Bcast(r, n, t, 0, c);
}
