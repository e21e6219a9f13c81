{-# LANGUAGE ForeignFunctionInterface #-}
{- | GPU (B200) back-end for hbayes' exact-inference hot path, over the C ABI of @libhbayes_cuda@
     (@include/hbayes_cuda.h@).

NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no GHC.  The module is written against
hbayes 0.5.2 (the reference) and the header above; it is the binding INTEGRATION.md describes.

It keeps the reference's names and argument shapes for the path it replaces:

  * 'createJunctionTree'  (Bayes/FactorElimination.hs:298-307)
  * 'changeEvidence'      (Bayes/FactorElimination/JTree.hs:454-466)
  * 'posterior'           (Bayes/FactorElimination.hs:314-327)
  * 'priorMarginal', 'posteriorMarginal' (Bayes/VariableElimination.hs:116-147)
  * 'changeFactor' on a tree (Bayes/FactorElimination/JTree.hs:107-115), 'learnEM' (Bayes/EM.hs:36-45)
  * 'mpe' (Bayes/VariableElimination.hs:101-114)

Networks are still built with the reference's own model layer ('Bayes.BayesianNetwork', or
'Bayes.ImportExport.HuginNet.importBayesianGraph' + 'runBN'); only inference moves to the GPU.
The pure record @JunctionTree CPT@ becomes the opaque 'GpuJunctionTree' (a 'ForeignPtr' with
@hb_jt_free@ as finalizer); results come back as ordinary 'CPT's built with 'factorWithVariables'.
All imports are @safe@: calls last milliseconds to seconds.
-}
module Bayes.Cuda
  ( GpuJunctionTree
  , TriangulationOrder(..)
  , createJunctionTree
  , changeEvidence
  , posterior
  , changeFactor
  , learnEM
  , mpe
  , posteriorsBatch
  , specialise
  , priorMarginal
  , posteriorMarginal
  ) where

import           Bayes                         (SBN, allVertexValues, fmapWithVertex)
import           Bayes.Factor                  (DV (..), DVI, Factor (..), Vertex (..), instantiationVariable)
import           Bayes.Factor.CPT              (CPT)
import           Bayes.PrivateTypes            (DVI (..))
import           Control.Monad                 (forM, when)
import           Data.Maybe                    (fromJust)
import           Data.Int                      (Int32, Int64, Int8)
import           Foreign
import           Foreign.C.String
import           Foreign.C.Types
import           System.IO.Unsafe              (unsafePerformIO)

data HbNet
data HbJt
data HbVe

foreign import ccall safe "hb_last_error"  c_last_error :: IO CString
foreign import ccall safe "hb_net_create"
  c_net_create :: Int32 -> Ptr Int32 -> Ptr Int64 -> Ptr Int32 -> Ptr Int64 -> Ptr Double -> Ptr (Ptr HbNet) -> IO CInt
foreign import ccall safe "&hb_net_free"   p_net_free :: FunPtr (Ptr HbNet -> IO ())
foreign import ccall safe "hb_jt_compile"
  c_jt_compile :: Ptr HbNet -> Int32 -> Ptr Int32 -> Int32 -> Ptr (Ptr HbJt) -> IO CInt
foreign import ccall safe "hb_jt_compile_with_order"
  c_jt_compile_with_order :: Ptr HbNet -> Ptr Int32 -> Ptr Int32 -> Int32 -> Ptr (Ptr HbJt) -> IO CInt
foreign import ccall safe "&hb_jt_free"    p_jt_free :: FunPtr (Ptr HbJt -> IO ())
foreign import ccall safe "hb_jt_set_evidence"
  c_jt_set_evidence :: Ptr HbJt -> Int32 -> Ptr Int32 -> Ptr Int32 -> IO CInt
foreign import ccall safe "hb_jt_posterior"
  c_jt_posterior :: Ptr HbJt -> Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Double -> Int64 -> Ptr Int64 -> IO CInt
foreign import ccall safe "hb_jt_posteriors_batch"
  c_jt_posteriors_batch :: Ptr HbJt -> Int64 -> Ptr Int8 -> Ptr Double -> Int32 -> IO CInt
foreign import ccall safe "hb_jt_specialise"
  c_jt_specialise :: Ptr HbJt -> Int32 -> Ptr Int32 -> Ptr Int32 -> IO CInt
foreign import ccall safe "hb_host_alloc"  c_host_alloc :: Int64 -> Ptr (Ptr a) -> IO CInt
foreign import ccall safe "&hb_host_free"  p_host_free :: FunPtr (Ptr a -> IO ())
foreign import ccall safe "hb_jt_change_factor"
  c_jt_change_factor :: Ptr HbJt -> Int32 -> Ptr Int32 -> Ptr Double -> Ptr Int32 -> IO CInt
foreign import ccall safe "hb_jt_em_step"
  c_jt_em_step :: Ptr HbJt -> Int64 -> Ptr Int8 -> Ptr Double -> Ptr Int32 -> Int32 -> IO CInt
foreign import ccall safe "hb_ve_compile"
  c_ve_compile :: Ptr HbNet -> Ptr Int32 -> Int32 -> Ptr Int32 -> Int32 -> Ptr Int32 -> Int32 -> Ptr (Ptr HbVe) -> IO CInt
foreign import ccall safe "hb_ve_compile_mpe"
  c_ve_compile_mpe :: Ptr HbNet -> Ptr Int32 -> Int32 -> Ptr Int32 -> Int32 -> Ptr Int32 -> Int32 -> Ptr (Ptr HbVe) -> IO CInt
foreign import ccall safe "hb_ve_mpe"
  c_ve_mpe :: Ptr HbVe -> Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Int8 -> Ptr Double -> Ptr Int32 -> Ptr Int32 -> IO CInt
foreign import ccall safe "&hb_ve_free"    p_ve_free :: FunPtr (Ptr HbVe -> IO ())
foreign import ccall safe "hb_ve_posterior"
  c_ve_posterior :: Ptr HbVe -> Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Double -> Int64 -> Ptr Int64 -> IO CInt

-- | How the triangulation picks vertices.  The reference takes an arbitrary closure
-- @UndirectedSG () f -> Vertex -> Vertex -> Ordering@; its metric is evaluated on the ORIGINAL moral
-- graph only (Bayes/FactorElimination.hs:291-292), so any closure reduces to a vertex order, which the
-- caller can compute in Haskell and pass with 'ExplicitOrder'.
data TriangulationOrder
  = WeightedEdges            -- ^ 'nodeComparisonForTriangulation' (Bayes/FactorElimination.hs:106-111)
  | NumberOfAddedEdges       -- ^ @compare `on` numberOfAddedEdges g@ (:71-78)
  | ExplicitOrder [Vertex]

-- | The calibrated tree on the GPU together with what is needed to rebuild 'CPT's on the way back.
data GpuJunctionTree = GpuJunctionTree
  { jtHandle :: ForeignPtr HbJt
  , jtDims   :: [Int]          -- ^ dimension of each vertex, index = vertex id
  }

check :: CInt -> IO ()
check 0 = return ()
check c = do
  msg <- c_last_error >>= peekCString
  ioError (userError ("libhbayes_cuda (" ++ show c ++ "): " ++ msg))

vertexId :: Vertex -> Int32
vertexId (Vertex i) = fromIntegral i

-- | Marshal the CPTs of a network.  'allVertexValues' is ascending vertex id (Bayes.hs:568); a CPT's
-- 'factorVariables' are [child, parents in ingoing order] (Bayes/Network.hs:69-79) and its values come out
-- in layout order through 'factorToList'.
withNet :: SBN CPT -> (Ptr HbNet -> [Int] -> IO a) -> IO a
withNet g k = do
  let cpts   = allVertexValues g
      scopes = map factorVariables cpts
      vals   = map factorToList cpts
      dims   = [ d | (DV _ d : _) <- scopes ]
      offs xs = scanl (+) 0 (map (fromIntegral . length) xs) :: [Int64]
  withArrayLen (map fromIntegral dims) $ \n pDims ->
    withArray (offs scopes) $ \pCptOff ->
    withArray [ vertexId v | sc <- scopes, DV v _ <- sc ] $ \pScope ->
    withArray (offs vals) $ \pValOff ->
    withArray (concat vals) $ \pVals ->
    alloca $ \pOut -> do
      c_net_create (fromIntegral n) pDims pCptOff pScope pValOff pVals pOut >>= check
      net <- peek pOut >>= newForeignPtr p_net_free
      withForeignPtr net (\p -> k p dims)

-- | 'createJunctionTree' on the GPU (device 0).  Pure in the reference; here the handle is created once
-- and never mutated by anything but 'changeEvidence', so 'unsafePerformIO' keeps the reference's type.
createJunctionTree :: TriangulationOrder -> SBN CPT -> GpuJunctionTree
createJunctionTree order g = unsafePerformIO $ withNet g $ \net dims ->
  with (0 :: Int32) $ \pDev -> alloca $ \pOut -> do
    case order of
      WeightedEdges      -> c_jt_compile net 0 pDev 1 pOut >>= check
      NumberOfAddedEdges -> c_jt_compile net 1 pDev 1 pOut >>= check
      ExplicitOrder vs   -> withArray (map vertexId vs) $ \pOrd -> c_jt_compile_with_order net pOrd pDev 1 pOut >>= check
    h <- peek pOut >>= newForeignPtr p_jt_free
    return (GpuJunctionTree h dims)

dviPair :: DVI -> (Int32, Int32)
dviPair i@(DVI _ s) = let DV v _ = instantiationVariable i in (vertexId v, fromIntegral s)

-- | 'changeEvidence': the list order is kept (it decides the reference's operation order, SURVEY A.9).
-- The handle carries the evidence state: like the reference's value it must not be shared between threads
-- that set different evidence.
changeEvidence :: [DVI] -> GpuJunctionTree -> GpuJunctionTree
changeEvidence e jt = unsafePerformIO $ withForeignPtr (jtHandle jt) $ \h -> do
  let (vs, ss) = unzip (map dviPair e)
  withArrayLen vs $ \n pV -> withArray ss $ \pS -> c_jt_set_evidence h (fromIntegral n) pV pS >>= check
  return jt

-- | 'posterior': 'Nothing' when no cluster holds all the variables (HB_E_NO_CLUSTER = -2).
posterior :: GpuJunctionTree -> [DV] -> Maybe CPT
posterior jt dvs = unsafePerformIO $ withForeignPtr (jtHandle jt) $ \h -> do
  let n    = length dvs
      size = product [ d | DV _ d <- dvs ]
  withArray [ vertexId v | DV v _ <- dvs ] $ \pQ ->
    allocaArray n $ \pScope -> allocaArray size $ \pOut -> alloca $ \pN -> do
      rc <- c_jt_posterior h (fromIntegral n) pQ pScope pOut (fromIntegral size) pN
      if rc == (-2) then return Nothing else do
        check rc
        scope <- peekArray n pScope
        vals  <- peekArray size pOut
        let dvOf i = DV (Vertex (fromIntegral i)) (jtDims jt !! fromIntegral i)
        return (factorWithVariables (map dvOf scope) vals)

-- | 'changeFactor' of @instance FactorContainer (JTree Cluster)@: the CPT with the same variable LIST gets the new
-- values, the evidence is kept, the tree recalibrated; on the GPU the compiled schedules stay (a potential upload).
-- A factor that matches no CPT leaves the tree unchanged, as in the reference.
changeFactor :: CPT -> GpuJunctionTree -> GpuJunctionTree
changeFactor f jt = unsafePerformIO $ withForeignPtr (jtHandle jt) $ \h -> do
  let scope = factorVariables f
  withArrayLen [ vertexId v | DV v _ <- scope ] $ \n pS -> withArray (factorToList f) $ \pV ->
    c_jt_change_factor h (fromIntegral n) pS pV nullPtr >>= check
  return jt

-- | 'learnEM samples g' (Bayes/EM.hs:36-45): one expectation/maximisation step, all samples in one GPU batch.
-- The learnt tables come back with the variable order the reference's cptDivide gives them
-- ([parents in the order of their posterior .., child]) and replace the vertex values of @g@.
learnEM :: [[DVI]] -> SBN CPT -> SBN CPT
learnEM samples g = unsafePerformIO $ do
  let jt     = createJunctionTree WeightedEdges g
      dims   = jtDims jt
      nVars  = length dims
      scopes = map factorVariables (allVertexValues g)
      sizes  = [ product [ d | DV _ d <- sc ] | sc <- scopes ]
      row s  = [ maybe (-1) fromIntegral (lookup v [ (fromIntegral a, b) | (a, b) <- map dviPair s ]) | v <- [0 .. nVars - 1] ] :: [Int8]
  withForeignPtr (jtHandle jt) $ \h ->
    withArray (concatMap row samples) $ \pEv ->
    allocaArray (sum sizes) $ \pOut -> allocaArray (sum (map length scopes)) $ \pScopes -> do
      c_jt_em_step h (fromIntegral (length samples)) pEv pOut pScopes 0 >>= check
      vals <- peekArray (sum sizes) pOut
      scs  <- peekArray (sum (map length scopes)) pScopes
      let dvOf i = DV (Vertex (fromIntegral i)) (dims !! fromIntegral i)
          cut ns xs = case ns of { [] -> []; (k : ks) -> let (a, b) = splitAt k xs in a : cut ks b }
          tables = [ fromJust (factorWithVariables (map dvOf sc) vs)
                   | (sc, vs) <- zip (cut (map length scopes) scs) (cut sizes vals) ] :: [CPT]
      return (fmapWithVertex (\(Vertex v) _ -> tables !! v) g)

-- | New, additive: one junction-tree query per evidence row (-1 = unobserved); the result holds, per
-- row, the posterior of every variable in ascending vertex id.
--
-- Both matrices live in page-locked memory from @hb_host_alloc@ (released by the 'ForeignPtr' finalizer
-- @hb_host_free@): GHC's own @mallocForeignPtrBytes@ / @allocaArray@ memory is pageable, and the library moves a
-- pageable 2^20-row batch in 53 ms where a page-locked one takes 17 ms.
posteriorsBatch :: GpuJunctionTree -> [[Int8]] -> IO [[Double]]
posteriorsBatch jt rows = withForeignPtr (jtHandle jt) $ \h -> do
  let nRows = length rows
      nVars = length (jtDims jt)
      width = sum (jtDims jt)
  ev  <- pinned (nRows * nVars)                          :: IO (ForeignPtr Int8)
  out <- pinned (nRows * width * sizeOf (0 :: Double))   :: IO (ForeignPtr Double)
  withForeignPtr ev $ \pEv -> withForeignPtr out $ \pOut -> do
    pokeArray pEv (concat rows)
    c_jt_posteriors_batch h (fromIntegral nRows) pEv pOut 0 >>= check
    flat <- peekArray (nRows * width) pOut
    return (chunks width flat)
 where
  chunks _ [] = []
  chunks k xs = let (a, b) = splitAt k xs in a : chunks k b
  pinned bytes = alloca $ \pp -> do
    c_host_alloc (fromIntegral (max 1 bytes)) pp >>= check
    peek pp >>= newForeignPtr p_host_free

-- | New, additive: build the kernels generated for @changeEvidence@ on these variables now (NVRTC, or the cubin cache /
-- a saved handle) rather than in the background of the first large batch.  Returns how many schedule steps they cover.
specialise :: GpuJunctionTree -> [Vertex] -> IO Int
specialise jt observed = withForeignPtr (jtHandle jt) $ \h ->
  withArrayLen [ vertexId v | v <- observed ] $ \n pV -> alloca $ \pN -> do
    c_jt_specialise h (fromIntegral n) pV pN >>= check
    fromIntegral `fmap` peek pN

-- | 'posteriorMarginal g p r assignment' (result variable order follows the reference, SURVEY A.5).
posteriorMarginal :: SBN CPT -> [DV] -> [DV] -> [DVI] -> CPT
posteriorMarginal g p r assignment = unsafePerformIO $ withNet g $ \net dims ->
  withArrayLen [ vertexId v | DV v _ <- p ] $ \np pP ->
  withArrayLen [ vertexId v | DV v _ <- r ] $ \nr pR ->
  with (0 :: Int32) $ \pDev -> alloca $ \pOut -> do
    c_ve_compile net pP (fromIntegral np) pR (fromIntegral nr) pDev 1 pOut >>= check
    ve <- peek pOut >>= newForeignPtr p_ve_free
    let (vs, ss) = unzip (map dviPair assignment)
        size     = product [ d | DV _ d <- r ]
    withForeignPtr ve $ \h -> withArrayLen vs $ \n pV -> withArray ss $ \pS ->
      allocaArray nr $ \pScope -> allocaArray size $ \pVals -> alloca $ \pN -> do
        c_ve_posterior h (fromIntegral n) pV pS pScope pVals (fromIntegral size) pN >>= check
        scope <- peekArray nr pScope
        vals  <- peekArray size pVals
        let dvOf i = DV (Vertex (fromIntegral i)) (dims !! fromIntegral i)
        maybe (ioError (userError "factorWithVariables")) return (factorWithVariables (map dvOf scope) vals)

-- | 'mpe g someP someR assignment': the variables of @someP@ are summed out (the evidence variables belong there),
-- those of @someR@ maximised.  The result is the reference's @[DVISet]@: one instantiation, in the reference's order
-- (the library reports that order; ties go to the last maximal state as with 'maximumBy').
mpe :: SBN CPT -> [DV] -> [DV] -> [DVI] -> [[DVI]]
mpe g p r assignment = unsafePerformIO $ withNet g $ \net dims ->
  withArrayLen [ vertexId v | DV v _ <- p ] $ \np pP ->
  withArrayLen [ vertexId v | DV v _ <- r ] $ \nr pR ->
  with (0 :: Int32) $ \pDev -> alloca $ \pOut -> do
    c_ve_compile_mpe net pP (fromIntegral np) pR (fromIntegral nr) pDev 1 pOut >>= check
    ve <- peek pOut >>= newForeignPtr p_ve_free
    let (vs, ss) = unzip (map dviPair assignment)
    withForeignPtr ve $ \h -> withArrayLen vs $ \n pV -> withArray ss $ \pS ->
      allocaArray nr $ \pStates -> alloca $ \pValue -> allocaArray nr $ \pOrder -> alloca $ \pN -> do
        c_ve_mpe h (fromIntegral n) pV pS pStates pValue pOrder pN >>= check
        states <- peekArray nr pStates
        k      <- peek pN
        order  <- peekArray (fromIntegral k) pOrder
        let byVar = zip [ vertexId v | DV v _ <- r ] states
            dvi i = DVI (DV (Vertex (fromIntegral i)) (dims !! fromIntegral i)) (maybe 0 fromIntegral (lookup i byVar))
        return (if null order then [] else [map dvi order])

-- | 'priorMarginal g ea eb' = 'posteriorMarginal' without evidence (Bayes/VariableElimination.hs:134-147).
priorMarginal :: SBN CPT -> [DV] -> [DV] -> CPT
priorMarginal g ea eb = posteriorMarginal g ea eb []
