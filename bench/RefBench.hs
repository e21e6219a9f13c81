{- | CPU baseline over the UNMODIFIED reference (alpheccar/hbayes 0.5.2) for the path this repository accelerates.

NOT COMPILED IN THIS REPOSITORY'S CI: there is no GHC in the build image or on the GPU boxes (DESIGN.md §7), so
the CPU figures reported by bench.py come from the C++ restatement of the reference algorithm (oracle/, kind
"port").  Anyone with GHC >= 7.10 and the dependencies of hbayes.cabal:147-168 can produce the true number:

  ghc -O2 -threaded -rtsopts -i/path/to/hbayes bench/RefBench.hs -o refbench
  ./refbench net.net evidence.txt +RTS -N

net.net        the network in the Hugin dialect of Bayes/ImportExport/HuginNet.hs
               (python -c "from hbayes_b200 import synth; synth.config_c2()[0].write_hugin('net.net')")
evidence.txt   one row per line, n_vars integers, -1 = unobserved
               (numpy.savetxt('evidence.txt', synth.config_c2_evidence(net, observed, rows), fmt='%d'))

One query = changeEvidence row + posterior [v] for every variable v (SURVEY 8d), exactly what one row of
hb_jt_posteriors_batch computes.  The library is single-threaded; rows are spread over capabilities with forkIO.
-}
module Main (main) where

import           Bayes                          (SBN, allVertexValues)
import           Bayes.BayesianNetwork          (runBN)
import           Bayes.Factor                   (DV (..), factorMainVariable, factorToList, (=:), setDVValue)
import           Bayes.Factor.CPT               (CPT)
import           Bayes.FactorElimination        (changeEvidence, createJunctionTree, nodeComparisonForTriangulation, posterior)
import           Bayes.ImportExport.HuginNet    (importBayesianGraph)
import           Control.Concurrent             (forkIO, getNumCapabilities)
import           Control.Concurrent.MVar
import           Control.Exception              (evaluate)
import           Control.Monad                  (forM, forM_)
import           Data.Maybe                     (fromJust)
import           Data.Time.Clock                (diffUTCTime, getCurrentTime)
import           System.Environment             (getArgs)

main :: IO ()
main = do
  [netFile, evFile] <- getArgs
  Just bn <- importBayesianGraph netFile
  let (_, g) = runBN bn :: ((), SBN CPT)
      vars   = map factorMainVariable (allVertexValues g)          -- ascending vertex id
      jt     = createJunctionTree nodeComparisonForTriangulation g
  rows <- (map (map read . words) . lines) `fmap` readFile evFile :: IO [[Int]]
  caps <- getNumCapabilities
  let chunks = splitInto caps rows
      query row =
        let ev  = [ setDVValue dv s | (dv, s) <- zip vars row, s >= 0 ]
            jt' = changeEvidence ev jt
        in sum [ sum (factorToList (fromJust (posterior jt' [dv]))) | dv <- vars ]   -- forces every posterior
  t0 <- getCurrentTime
  dones <- forM chunks $ \chunk -> do
    mv <- newEmptyMVar
    _ <- forkIO (evaluate (sum (map query chunk)) >>= putMVar mv)
    return mv
  total <- sum `fmap` mapM takeMVar dones
  t1 <- getCurrentTime
  let secs = realToFrac (diffUTCTime t1 t0) :: Double
  putStrLn ("{\"impl\": \"hbayes (GHC)\", \"rows\": " ++ show (length rows) ++ ", \"capabilities\": " ++ show caps
            ++ ", \"seconds\": " ++ show secs ++ ", \"queries_per_s\": " ++ show (fromIntegral (length rows) / secs)
            ++ ", \"checksum\": " ++ show total ++ "}")

splitInto :: Int -> [a] -> [[a]]
splitInto n xs = go xs
  where
    k = max 1 ((length xs + n - 1) `div` n)
    go [] = []
    go ys = let (a, b) = splitAt k ys in a : go b
