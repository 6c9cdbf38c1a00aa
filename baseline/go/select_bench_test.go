// NOT EXECUTED in this repository's build image (no Go toolchain).  A pure-function restatement of the
// reference's hot loops with the same shapes as the originals, for timing on a machine that has Go:
//
//	GetGPUTypes filter + sort.Slice + take 5      pkg/virtual_kubelet/runpod_client.go:465-509
//	updateAllPodStatuses diff predicate           pkg/virtual_kubelet/kubelet.go:870-873
//
// Tables follow SURVEY.md 8d (splitmix64, seed 0x52504B31); sizes are the BASELINE configs.
package rpkbaseline

import (
	"sort"
	"testing"
)

type gpuType struct {
	ID, DisplayName string
	MemoryInGb      int
	SecureCloud     bool
	SecurePrice     float64
	CommunityCloud  bool
	CommunityPrice  float64
}

type instanceInfo struct {
	Status       string
	PortsExposed bool
}

func splitmix64(x uint64) uint64 {
	z := x + 0x9E3779B97F4A7C15
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9
	z = (z ^ (z >> 27)) * 0x94D049BB133111EB
	return z ^ (z >> 31)
}

func r(col, row uint64) uint64 { return splitmix64(0x52504B31 + (col << 40) + row) }

var memSet = []int{8, 12, 16, 20, 24, 32, 40, 48, 80, 94, 141, 180, 192}
var reqMemSet = []int{2, 8, 16, 24, 40, 48, 80}

func makeOffers(g int) []gpuType {
	out := make([]gpuType, g)
	for i := range out {
		row := uint64(i)
		cents := int64(5 + r(3, row)%395)
		if r(4, row)%100 < 2 {
			cents = 0
		}
		out[i] = gpuType{
			ID:             "gpu",
			MemoryInGb:     memSet[r(0, row)%uint64(len(memSet))],
			SecureCloud:    r(1, row)%100 < 75,
			CommunityCloud: r(2, row)%100 < 50,
			SecurePrice:    float64(cents) / 100.0,
			CommunityPrice: float64(cents*6/10) / 100.0,
		}
	}
	return out
}

type podReq struct {
	minRAM    int
	cloudType string
}

func makePods(p int) []podReq {
	out := make([]podReq, p)
	for i := range out {
		row := uint64(i)
		mem := reqMemSet[r(17, row)%uint64(len(reqMemSet))]
		if r(16, row)%100 < 40 {
			mem = 16
		}
		cloud := "SECURE"
		if r(18, row)%100 >= 90 {
			cloud = "COMMUNITY"
		}
		out[i] = podReq{mem, cloud}
	}
	return out
}

// survivor is one offer that passed the filter.
type survivor struct {
	id    string
	price float64
}

// cheapestFive restates what Client.GetGPUTypes computes once the offer table is decoded
// (runpod_client.go:465-509): keep offers whose cloud flag is set for the requested cloud type and whose
// price lies strictly inside (0, maxPrice) and whose memory is at least minRAM; order them by price with
// sort.Slice (so ties fall wherever Go's pdqsort puts them); return at most five ids, never nil.
func cheapestFive(table []gpuType, minRAM int, maxPrice float64, cloud string) []string {
	keep := make([]survivor, 0, 16)
	for i := range table {
		t := &table[i]
		price, offered := 0.0, false
		switch cloud {
		case "SECURE":
			price, offered = t.SecurePrice, t.SecureCloud
		case "COMMUNITY":
			price, offered = t.CommunityPrice, t.CommunityCloud
		}
		if !offered || !(price > 0) || !(price < maxPrice) || t.MemoryInGb < minRAM {
			continue
		}
		keep = append(keep, survivor{t.ID, price})
	}
	sort.Slice(keep, func(a, b int) bool { return keep[a].price < keep[b].price })
	if len(keep) > 5 {
		keep = keep[:5]
	}
	ids := make([]string, len(keep))
	for i, k := range keep {
		ids[i] = k.id
	}
	return ids
}

func benchSelect(b *testing.B, p, g int) {
	types := makeOffers(g)
	pods := makePods(p)
	b.ResetTimer()
	for it := 0; it < b.N; it++ {
		for _, pod := range pods {
			_ = cheapestFive(types, pod.minRAM, 0.5, pod.cloudType)
		}
	}
	b.ReportMetric(float64(p)*float64(g)*float64(b.N)/b.Elapsed().Seconds(), "offer-scores/s")
}

func BenchmarkSelectC1_100x50(b *testing.B)    { benchSelect(b, 100, 50) }
func BenchmarkSelectC2_10kx1k(b *testing.B)    { benchSelect(b, 10_000, 1_000) }
func BenchmarkSelectC3_2kx10k(b *testing.B)    { benchSelect(b, 2_000, 10_000) }   // row sample of C3
func BenchmarkSelectC4_500x100k(b *testing.B)  { benchSelect(b, 500, 100_000) }    // row sample of C4

// the diff predicate of updateAllPodStatuses (kubelet.go:870-873) over n tracked pods
func BenchmarkStatusSweep100k(b *testing.B) {
	names := []string{"RUNNING", "STARTING", "EXITED", "TERMINATING", "TERMINATED", "NOT_FOUND"}
	n := 100_000
	prev := make([]instanceInfo, n)
	now := make([]instanceInfo, n)
	for i := range prev {
		prev[i] = instanceInfo{names[r(32, uint64(i))%6], r(34, uint64(i))%100 < 85}
		now[i] = prev[i]
		if r(44, uint64(i))%100 < 1 {
			now[i] = instanceInfo{names[r(45, uint64(i))%6], r(47, uint64(i))%100 < 85}
		}
	}
	b.ResetTimer()
	changed := 0
	for it := 0; it < b.N; it++ {
		for i := range now {
			statusChanged := now[i].Status != prev[i].Status
			portsExposureChanged := now[i].PortsExposed != prev[i].PortsExposed
			if statusChanged || portsExposureChanged {
				changed++
			}
		}
	}
	_ = changed
	b.ReportMetric(float64(n)*float64(b.N)/b.Elapsed().Seconds(), "pods/s")
}
