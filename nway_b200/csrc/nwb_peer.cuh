// Synchronisation between the GPUs of one node through flags in peer memory (cudaIpc-mapped buffers): what the shard
// mode's barrier and the table reassembly need of a collective library, without one.  One warp; lane r talks to rank r.
//   publish   (optional) a payload word into slot [rank] of every peer's payload array, then -- after a system-scope
//             fence -- the tag into slot [rank] of every peer's tag array (release);
//   wait      until every slot of the OWN tag array carries a tag >= this one (acquire); tags only grow;
//   collect   (optional) the peers' payload words into a device array and into mapped host memory.
// Everything a previous kernel of this stream stored into peer memory has been performed when this kernel starts, so the
// tag a peer sees also tells it that those stores have landed.  A wait gives up after timeout_ns (a peer that died would
// otherwise hang the GPU): it raises abort[0] and the host word err[0]; the kernels that follow look at abort.
#pragma once

struct PeerSync {
	char *peers[16];               // every rank's buffer base ([rank] = the own one)
	long long tag_off;             // byte offset of the tag array (16 words) in every buffer
	long long payload_off;         // byte offset of the payload array, -1 = none
	unsigned long long tag;
	long long payload;
	long long *collect_dev;        // [17]: payloads by rank, [16] = abort flag; null = none
	long long *collect_host;       // mapped host memory, same layout ([16] = error code); may be null
	unsigned long long timeout_ns;
	int world, rank;
};

__device__ __forceinline__ unsigned long long peer_globaltimer()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

__global__ void __launch_bounds__(32) k_peer_sync(PeerSync A)
{
	const int r = threadIdx.x;
	bool ok = true;
	if (r < A.world) {
		if (A.payload_off >= 0) {
			*reinterpret_cast<volatile long long *>(A.peers[r] + A.payload_off + 8 * A.rank) = A.payload;
			__threadfence_system();
		}
		unsigned long long *flag = reinterpret_cast<unsigned long long *>(A.peers[r] + A.tag_off + 8 * A.rank);
		__threadfence_system();
		asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"(A.tag) : "memory");
		const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(A.peers[A.rank] + A.tag_off + 8 * r);
		const unsigned long long t0 = peer_globaltimer();
		for (;;) {
			unsigned long long v;
			asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
			if (v >= A.tag) break;
			if (peer_globaltimer() - t0 > A.timeout_ns) { ok = false; break; }
			__nanosleep(100);
		}
	}
	const bool all_ok = __all_sync(0xffffffffu, ok);
	if (r < A.world && A.payload_off >= 0 && A.collect_dev) {
		const long long v = all_ok ? *reinterpret_cast<volatile const long long *>(A.peers[A.rank] + A.payload_off + 8 * r) : 0;
		A.collect_dev[r] = v;
		if (A.collect_host) A.collect_host[r] = v;
	}
	if (r == 0 && !all_ok) {
		if (A.collect_dev) A.collect_dev[16] = 1;
		if (A.collect_host) A.collect_host[16] = 1;
	}
}
