// Every collective of the bundled MPI stand-in (hysortk_b200/shim/mpi.h) on N ranks, with payloads larger than one staging
// slot, non-blocking all-to-all rounds with a barrier in between (the reference's pattern, kmerops.cpp:814-968), and
// "abort" as argv[1]: rank 1 calls MPI_Abort while the others wait in a barrier.
#include <mpi.h>
#include <cassert>
#include <cstdio>
#include <numeric>
#include <string>
#include <cstring>
#include <vector>
int main(int argc,char**argv){
  MPI_Init(&argc,&argv);
  int r,n; MPI_Comm_rank(MPI_COMM_WORLD,&r); MPI_Comm_size(MPI_COMM_WORLD,&n);
  if (argc > 1 && std::string(argv[1]) == "abort") { if (r == 1) MPI_Abort(MPI_COMM_WORLD, 7); MPI_Barrier(MPI_COMM_WORLD); MPI_Barrier(MPI_COMM_WORLD); MPI_Finalize(); return 0; }
  // bcast big (multi-round)
  std::vector<unsigned long long> big(3000000);
  if(r==0) for(size_t i=0;i<big.size();++i) big[i]=i*7+1;
  MPI_Bcast(big.data(),(int)big.size(),MPI_UNSIGNED_LONG_LONG,0,MPI_COMM_WORLD);
  for(size_t i=0;i<big.size();i+=997) assert(big[i]==i*7+1);
  int x=r+1,sum=0; MPI_Allreduce(&x,&sum,1,MPI_INT,MPI_SUM,MPI_COMM_WORLD); assert(sum==n*(n+1)/2);
  int mx=r; MPI_Allreduce(MPI_IN_PLACE,&mx,1,MPI_INT,MPI_MAX,MPI_COMM_WORLD); assert(mx==n-1);
  double d=r*1.5,dm=0; MPI_Reduce(&d,&dm,1,MPI_DOUBLE,MPI_MAX,0,MPI_COMM_WORLD); if(r==0) assert(dm==(n-1)*1.5);
  int ex=-5; int one=r+1; MPI_Exscan(&one,&ex,1,MPI_INT,MPI_SUM,MPI_COMM_WORLD); if(r>0) assert(ex==r*(r+1)/2); else assert(ex==-5);
  // alltoallv with variable sizes, large
  std::vector<int> sn(n),sd(n),rn(n),rd(n);
  for(int p=0;p<n;++p){ sn[p]=100000*(r+1)+p; rn[p]=100000*(p+1)+r; }
  std::exclusive_scan(sn.begin(),sn.end(),sd.begin(),0); std::exclusive_scan(rn.begin(),rn.end(),rd.begin(),0);
  std::vector<unsigned long long> sb(sd[n-1]+sn[n-1]), rb(rd[n-1]+rn[n-1]);
  for(int p=0;p<n;++p) for(int i=0;i<sn[p];++i) sb[sd[p]+i]=((unsigned long long)r<<40)|((unsigned long long)p<<32)|i;
  MPI_Alltoallv(sb.data(),sn.data(),sd.data(),MPI_UNSIGNED_LONG_LONG,rb.data(),rn.data(),rd.data(),MPI_UNSIGNED_LONG_LONG,MPI_COMM_WORLD);
  for(int p=0;p<n;++p) for(int i=0;i<rn[p];i+=13) assert(rb[rd[p]+i]==(((unsigned long long)p<<40)|((unsigned long long)r<<32)|i));
  // ialltoall rounds
  const int B=80000; std::vector<char> s1(B*n), r1(B*n);
  for(int it=0;it<20;++it){ for(int p=0;p<n;++p) memset(&s1[p*B], (r*16+p+it)&0xFF, B); MPI_Request q; MPI_Ialltoall(s1.data(),B,MPI_BYTE,r1.data(),B,MPI_BYTE,MPI_COMM_WORLD,&q); if (it % 3 == 0) MPI_Barrier(MPI_COMM_WORLD); MPI_Wait(&q,MPI_STATUS_IGNORE);
    for(int p=0;p<n;++p){ assert((unsigned char)r1[p*B]==((p*16+r+it)&0xFF)); assert((unsigned char)r1[p*B+B-1]==((p*16+r+it)&0xFF)); } }
  // gather/gatherv/scatterv
  int g=r*3; std::vector<int> all(n); MPI_Gather(&g,1,MPI_INT,all.data(),1,MPI_INT,0,MPI_COMM_WORLD); if(r==0) for(int p=0;p<n;++p) assert(all[p]==p*3);
  std::vector<char> mine(r+1,'a'+r); std::vector<int> cn(n),cd(n); for(int p=0;p<n;++p){cn[p]=p+1;} std::exclusive_scan(cn.begin(),cn.end(),cd.begin(),0);
  std::vector<char> cat(cd[n-1]+cn[n-1]); MPI_Gatherv(mine.data(),r+1,MPI_CHAR,cat.data(),cn.data(),cd.data(),MPI_CHAR,0,MPI_COMM_WORLD);
  if(r==0) for(int p=0;p<n;++p) for(int i=0;i<cn[p];++i) assert(cat[cd[p]+i]=='a'+p);
  MPI_Datatype t3; MPI_Type_contiguous(3,MPI_UNSIGNED_LONG_LONG,&t3); MPI_Type_commit(&t3);
  std::vector<unsigned long long> root; if(r==0){ root.resize(3*(cd[n-1]+cn[n-1])); for(size_t i=0;i<root.size();++i) root[i]=i; }
  std::vector<unsigned long long> me(3*cn[r]); MPI_Scatterv(root.data(),cn.data(),cd.data(),t3,me.data(),cn[r],t3,0,MPI_COMM_WORLD);
  for(int i=0;i<3*cn[r];++i) assert(me[i]==(unsigned long long)(3*cd[r]+i));
  MPI_Barrier(MPI_COMM_WORLD);
  if(r==0) printf("shim ok n=%d\n",n);
  MPI_Finalize(); return 0; }
