// CPU check of the half-warp scheduler the dual-table layout uses (sched16.cuh): every lane keeps its
// entries, and every step touches 16 distinct bank classes whenever the class loads can be balanced.
#include "sched16.cuh"
#ifndef TNL
#define TNL 16
#endif
#include <vector>
#include <random>
#include <cstdio>
#include <cstring>
#include <algorithm>
struct Mem { std::vector<unsigned char> b; std::vector<uint32_t> w; unsigned char& B(int i){return b[i];} uint32_t& Wd(int i){return w[i];} };
struct Out { std::vector<unsigned char> c; int W; void operator()(int l,int t,unsigned char v){ c[l*W+t]=v; } };
int main(int argc,char**argv){
  std::mt19937 rng(1);
  int trials = 2000; long totalsteps=0,totalwaves=0, badcnt=0, ovf=0; long ms[4]={0},mw[4]={0},mo[4]={0};
  for(int tr=0;tr<trials;++tr){
    int p = 784, m = 78; int mode = tr%4;
    int W; std::vector<std::vector<int>> rows(TNL);
    if(mode==0){ W=78; for(auto&r:rows){ std::vector<int> all(p); for(int i=0;i<p;++i)all[i]=i; std::shuffle(all.begin(),all.end(),rng); r.assign(all.begin(),all.begin()+m);} }
    else if(mode==1){ p=1024; W=52; for(auto&r:rows){ std::vector<int> all(p); for(int i=0;i<p;++i)all[i]=i; std::shuffle(all.begin(),all.end(),rng); r.assign(all.begin(),all.begin()+51);} }
    else if(mode==2){ p=200; int mx=0; for(auto&r:rows){ int len=rng()%60; std::vector<int> all(p); for(int i=0;i<p;++i)all[i]=i; std::shuffle(all.begin(),all.end(),rng); r.assign(all.begin(),all.begin()+len); mx=std::max(mx,len);} W=(mx+1)&~1; if(W==0) W=2; }
    else { p=64; W=0; int mx=0; for(auto&r:rows){ int len=rng()%20; for(int i=0;i<len;++i) r.push_back(TNL*(rng()%4) + (rng()%3)); std::sort(r.begin(),r.end()); r.erase(std::unique(r.begin(),r.end()),r.end()); mx=std::max(mx,(int)r.size()); } W=(mx+1)&~1; if(W==0)W=2; }
    int wmax = W + (tr%3)*2;
    Mem M; M.b.assign(skm_sched_bytes(TNL,wmax),0xAB); M.w.assign(skm_sched_words(TNL,wmax),0xdeadbeef);
    int cnt[TNL][TNL]; memset(cnt,0,sizeof cnt);
    for(int l=0;l<TNL;++l){ for(int r:rows[l]) cnt[l][r&(TNL-1)]++; }
    for(int l=0;l<TNL;++l) for(int g=0;g<TNL;++g) M.b[l*TNL+g]=cnt[l][g];
    Out O; O.W=W; O.c.assign(TNL*W,0x77);
    #ifdef USE_BVN
    M.b.assign(skm_bvn_bytes(TNL),0xAB); for(int l=0;l<TNL;++l) for(int g=0;g<TNL;++g) M.b[l*TNL+g]=cnt[l][g];
    int o = skm_sched_bvn<TNL>(M,W,O); ovf+=o;
#else
    int o = skm_sched<TNL>(M,W,wmax,O); ovf+=o;
#endif
    // verify: per lane multiset of groups matches cnt; per step classes distinct
    int got[TNL][TNL]; memset(got,0,sizeof got);
    for(int t=0;t<W;++t){ int cls[TNL]; int used[TNL]={0};
      for(int l=0;l<TNL;++l){ unsigned char cd=O.c[l*W+t]; if(cd==0x77){badcnt++;} int c; if(cd&0x80) c=cd&(TNL-1); else { int g=cd&15; got[l][g]++; c=(g+((cd>>4)&1))&(TNL-1); } cls[l]=c; used[c]++; }
      int mx=0; for(int c=0;c<TNL;++c) mx=std::max(mx,used[c]); totalsteps++; totalwaves+=mx; ms[mode]++; mw[mode]+=mx; }
    mo[mode]+=o;
    for(int l=0;l<TNL;++l) for(int g=0;g<TNL;++g) if(got[l][g]!=cnt[l][g]) { badcnt++; }
  }
  printf("trials %d steps %ld waves %ld ratio %.4f bad %ld overflow %ld\n",trials,totalsteps,totalwaves,(double)totalwaves/totalsteps,badcnt,ovf);
  for(int i=0;i<4;++i) printf("mode %d ratio %.4f overflow %ld\n",i,(double)mw[i]/ms[i],mo[i]);
  // modes 0-2 (uniform / ragged random rows) must be conflict-free; mode 3 is adversarial (3 classes only)
  for(int i=0;i<3;++i) if(mw[i]!=ms[i]||mo[i]) return 2;
  return badcnt?1:0;
}
